// me_hme.cuh -- HME window arithmetic shared by me_picture.cu (core path) and me_b64.cu (complete driver): hme_level_0/1/2 of
// Source/Lib/Codec/motion_estimation.c:820-1113 restated as "fill in the search item" + "scale the result back".
#pragma once
#include "common.cuh"
#include "sad_small.cuh"
#include "../../include/svt_b200.h"

namespace b200 {

// ---- HME window arithmetic (shared by the three levels) ---------------------------------------
struct HmeSide {  // kept per item between prepare and finish
    int16_t origin_x, origin_y;
};

__device__ __forceinline__ void hme_clip(int16_t org, int16_t& origin, int16_t& sa, int16_t pad, int16_t pic, bool round8) {
    if ((int16_t)(org + origin) < -pad) {
        origin = (int16_t)(-pad - org);
        sa     = (int16_t)(sa - (-pad - (org + origin)));  // (sic) evaluates to sa: origin was just moved
    }
    if ((int16_t)(org + origin) > (int16_t)(pic - 1)) origin = (int16_t)(origin - ((org + origin) - (pic - 1)));
    if ((int16_t)(org + origin + sa) > pic) {
        const int16_t v = (int16_t)(sa - ((org + origin + sa) - pic));
        sa = v > 1 ? v : (int16_t)1;
    }
    if (round8) sa = (sa < 8) ? sa : (int16_t)(sa & ~0x07);
}

// One HME search of region (sr_w, sr_h) of block (bx, by) at `level` (0: 1/16 picture, 1: 1/4, 2: full),
// centred on the previous level's result (prev_x, prev_y; unused at level 0): hme_level_0/1/2 of
// motion_estimation.c restated as "fill in the search item".
__device__ __forceinline__ void hme_make_item(const SvtB200MePicture& cur, const SvtB200MePicture& rp, const SvtB200MeParams& p, int level,
                                              int sr_w, int sr_h, int bx, int by, int16_t prev_x, int16_t prev_y,
                                              SvtB200SadSearchItem& it_out, HmeSide& side_out) {
    const int shift = 2 - level;
    const int full_x = bx * 64, full_y = by * 64;
    const int blk_w = min(64, cur.width[2] - full_x) >> shift, blk_h = min(64, cur.height[2] - full_y) >> shift;
    const int16_t org_x = (int16_t)(full_x >> shift), org_y = (int16_t)(full_y >> shift);
    int16_t sa_w, sa_h, ox, oy, pad_w, pad_h;
    if (level == 0) {
        sa_w = (int16_t)((p.hme_l0_sa_w + 7) & ~0x07);
        sa_h = (int16_t)p.hme_l0_sa_h;
        ox   = (int16_t)(-(int16_t)((sa_w * 2) >> 1) + sa_w * sr_w);
        oy   = (int16_t)(-(int16_t)((sa_h * 2) >> 1) + sa_h * sr_h);
        pad_w = (int16_t)(rp.org_x[0] - 1);
        pad_h = (int16_t)(rp.org_y[0] - 1);
    } else if (level == 1) {
        sa_w = (int16_t)((p.hme_l1_sa_w + 7) & ~0x07);
        sa_h = (int16_t)p.hme_l1_sa_h;
        ox   = (int16_t)(-(sa_w >> 1) + (prev_x >> 1));
        oy   = (int16_t)(-(sa_h >> 1) + (prev_y >> 1));
        pad_w = (int16_t)(rp.org_x[1] - 1);
        pad_h = (int16_t)(rp.org_y[1] - 1);
    } else {
        sa_w = (int16_t)((p.hme_l2_sa_w + 7) & ~0x07);
        sa_h = (int16_t)p.hme_l2_sa_h;
        ox   = (int16_t)(-(sa_w >> 1) + prev_x);
        oy   = (int16_t)(-(sa_h >> 1) + prev_y);
        pad_w = pad_h = 63;
    }
    hme_clip(org_x, ox, sa_w, pad_w, (int16_t)rp.width[level], true);
    hme_clip(org_y, oy, sa_h, pad_h, (int16_t)rp.height[level], false);
    const int sub = p.hme_sub_sad ? 1 : 0;
    SvtB200SadSearchItem it;
    it.src_off    = (uint64_t)(uintptr_t)(cur.plane[level] + (size_t)(cur.org_y[level] + org_y) * cur.stride[level] + cur.org_x[level] + org_x);
    it.ref_off    = (uint64_t)(uintptr_t)(rp.plane[level] + (size_t)(rp.org_y[level] + org_y + oy) * rp.stride[level] + rp.org_x[level] + org_x + ox);
    it.src_stride = (uint32_t)(cur.stride[level] << sub);
    it.ref_stride = (uint32_t)(rp.stride[level] << sub);
    it.ref_step   = (uint32_t)rp.stride[level];
    it.block_w    = (uint16_t)blk_w;
    it.block_h    = (uint16_t)(blk_h >> sub);
    it.sa_w       = sa_w;
    it.sa_h       = sa_h;
    it.skip_search_line = 0;
    it.reserved   = 0;
    it_out   = it;
    side_out = HmeSide{ox, oy};
}

__device__ __forceinline__ void hme_finish_one(const SvtB200SadSearchResult& q, const HmeSide& sd, const SvtB200MeParams& p, int level,
                                               int16_t& out_x, int16_t& out_y, uint64_t& out_sad) {
    const int mul = level == 0 ? 4 : (level == 1 ? 2 : 1);
    uint64_t  sad = q.best_sad;
    if (p.hme_sub_sad) sad *= 2;
    out_sad = sad;
    out_x   = (int16_t)((int16_t)(q.x + sd.origin_x) * mul);
    out_y   = (int16_t)((int16_t)(q.y + sd.origin_y) * mul);
}


}  // namespace b200
