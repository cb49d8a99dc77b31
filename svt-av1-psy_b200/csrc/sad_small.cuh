// sad_small.cuh -- warp-level full search over a SMALL search area (shared by sad.cu and me_picture.cu).
//
// Small search areas (the 8x3 / 16x4 HME and ME refinements: a few dozen positions, blocks up to
// 64x64): the work is a few hundred VABSDIFF4 per lane, so staging through shared memory and CTA
// barriers would be the whole cost.  One WARP per search, no shared memory: lane = (x mod 8, row slice
// of 4); a lane walks its rows with two sliding funnel-shift windows over aligned words that come
// straight from L1 (the 8 x-lanes of a slice read the same sectors, the source word is a broadcast)
// and evaluates up to four search rows y at once against every source word; the 4 slices are added
// with two shuffles and the raster-order first minimum (compute_sad_c.c:58-101: strict '<') is the
// minimum of the 64-bit key (sad<<32 | y<<16 | x).
#pragma once
#include "common.cuh"
#include "../../include/svt_b200.h"

namespace b200 {

constexpr int kSmallSearchMaxPos = 256;  // searches with at most this many positions take the warp path

// aligned-word view of `n` bytes at p: word(j) = bytes [4j, 4j+4) of the run; only words holding a valid byte are read
struct ByteRun {
    const uint32_t* w;
    int shift, last;
    __device__ __forceinline__ ByteRun(const uint8_t* p, int n) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        w     = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        shift = (int)(a & 3) * 8;
        last  = (int)(((a & 3) + n - 1) >> 2);
    }
    __device__ __forceinline__ uint32_t raw(int j) const { return __ldg(w + (j < last ? j : last)); }
};


__device__ __forceinline__ SvtB200SadSearchResult sad_key_to_result(unsigned long long best) {
    SvtB200SadSearchResult r;
    if (best == ~0ull) {
        r.best_sad = 0xffffffu;
        r.x = r.y = -1;
    } else {
        r.best_sad = (uint32_t)(best >> 32);
        r.y        = (int16_t)((best >> 16) & 0xffff);
        r.x        = (int16_t)(best & 0xffff);
    }
    return r;
}

// all 32 lanes call; every lane returns the winning key (~0ull when no position qualifies)
__device__ __forceinline__ unsigned long long sad_search_warp(const uint8_t* __restrict__ src0, const uint8_t* __restrict__ ref0,
                                                              const SvtB200SadSearchItem& item, int lane) {
    const int bw = item.block_w, bh = item.block_h, sa_w = item.sa_w, sa_h = item.sa_h;
    unsigned long long best = ~0ull;
    if (sa_w <= 0 || sa_h <= 0 || bw <= 0 || bh <= 0) return best;
    const bool     skip = (bw == 16 && bh <= 16 && item.skip_search_line);
    const int      xs = lane & 7, slice = lane >> 3;
    const int      nw = (bw + 3) >> 2, tail = bw & 3;
    const uint32_t tailmask = tail ? ((1u << (tail * 8)) - 1u) : 0xffffffffu;
    constexpr int  YG = 4;  // search rows evaluated per pass over the source block (the loop body below is written for 4)
    for (int y0 = 0; y0 < sa_h; y0 += YG) {
        for (int x0 = 0; x0 < sa_w; x0 += 8) {
            const bool valid = x0 + xs < sa_w;
            const int  x = valid ? x0 + xs : sa_w - 1;  // idle lanes shadow the last column (stays inside the window)
            uint32_t   acc[YG];
#pragma unroll
            for (int k = 0; k < YG; k++) acc[k] = 0;
            for (int r = slice; r < bh; r += 4) {
                const ByteRun  S(src0 + (size_t)r * item.src_stride, bw);
                const uint8_t* rrow = ref0 + (size_t)r * item.ref_stride + x;
                // rows past the area repeat the last one; their sums are dropped below
                const ByteRun R0(rrow + (size_t)min(y0 + 0, sa_h - 1) * item.ref_step, bw), R1(rrow + (size_t)min(y0 + 1, sa_h - 1) * item.ref_step, bw),
                    R2(rrow + (size_t)min(y0 + 2, sa_h - 1) * item.ref_step, bw), R3(rrow + (size_t)min(y0 + 3, sa_h - 1) * item.ref_step, bw);
                uint32_t slo = S.raw(0), r0 = R0.raw(0), r1 = R1.raw(0), r2 = R2.raw(0), r3 = R3.raw(0);
                // all words but the last: the next aligned word always holds valid bytes, nothing to mask
#pragma unroll 4
                for (int j = 0; j < nw - 1; j++) {
                    const uint32_t shi = __ldg(S.w + j + 1), h0 = __ldg(R0.w + j + 1), h1 = __ldg(R1.w + j + 1), h2 = __ldg(R2.w + j + 1),
                                   h3 = __ldg(R3.w + j + 1);
                    const uint32_t sv = __funnelshift_r(slo, shi, S.shift);  // one source word against four search rows
                    acc[0] = __vsadu4(sv, __funnelshift_r(r0, h0, R0.shift)) + acc[0];
                    acc[1] = __vsadu4(sv, __funnelshift_r(r1, h1, R1.shift)) + acc[1];
                    acc[2] = __vsadu4(sv, __funnelshift_r(r2, h2, R2.shift)) + acc[2];
                    acc[3] = __vsadu4(sv, __funnelshift_r(r3, h3, R3.shift)) + acc[3];
                    slo = shi;
                    r0  = h0;
                    r1  = h1;
                    r2  = h2;
                    r3  = h3;
                }
                {  // last word: the following aligned word may lie past the row (clamped index), bytes past bw are masked
                    const uint32_t shi = S.raw(nw), h0 = R0.raw(nw), h1 = R1.raw(nw), h2 = R2.raw(nw), h3 = R3.raw(nw);
                    const uint32_t sv = __funnelshift_r(slo, shi, S.shift) & tailmask;
                    acc[0] = __vsadu4(sv, __funnelshift_r(r0, h0, R0.shift) & tailmask) + acc[0];
                    acc[1] = __vsadu4(sv, __funnelshift_r(r1, h1, R1.shift) & tailmask) + acc[1];
                    acc[2] = __vsadu4(sv, __funnelshift_r(r2, h2, R2.shift) & tailmask) + acc[2];
                    acc[3] = __vsadu4(sv, __funnelshift_r(r3, h3, R3.shift) & tailmask) + acc[3];
                }
            }
#pragma unroll
            for (int k = 0; k < YG; k++) {
                uint32_t a = acc[k];
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                a += __shfl_xor_sync(0xffffffffu, a, 16);
                const int y = y0 + k;
                if (valid && y < sa_h && !(skip && ((y & 1) == 0)) && a < 0xffffffu) {
                    const unsigned long long key = ((unsigned long long)a << 32) | ((unsigned long long)(uint32_t)y << 16) | (unsigned long long)(uint32_t)x;
                    best = key < best ? key : best;
                }
            }
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {  // the 4 slices hold identical keys; reduce over the 8 x-lanes
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    return best;
}

}  // namespace b200
