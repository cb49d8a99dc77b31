/* oracle/ref_driver.c -- TEST / BASELINE INFRASTRUCTURE, compiled INTO oracle/_ref/libsvtav1_ref.so.
 *
 * Whole-picture drivers that run the UNMODIFIED reference kernels (through the reference's own
 * dispatch pointers) over the same work lists the B200 T2 entry points take, so that
 *   (1) the T2 paths can be checked against the reference at picture scale, and
 *   (2) bench.py --impl reference / cpu_baseline can time the reference's CPU path (C tier, or the
 *       intrinsics-only AVX2 tier) on all host cores (OpenMP) for the same workload.
 * The driver arithmetic restates the reference's process-level loops (cited per function); every
 * pixel-level computation is done by reference code.  Nothing here is used by the product.
 *
 * Not available in this build of the AVX2 tier (needs NASM): the dav1d inverse transforms -- the
 * inverse-transform leg always runs the reference C kernels (stated in DESIGN.md / bench output). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "definitions.h"
#include "aom_dsp_rtcd.h"
#include "common_dsp_rtcd.h"
#include "cdef.h"
#include "restoration.h"
#include "convolve.h"


/* ---- minimal pthread parallel-for (no OpenMP runtime in this image) ------------------------------- */
typedef void (*ParBody)(int i);
static struct { ParBody body; int n, chunk; volatile int next; } g_par;
static int g_threads = 0;
static void* par_worker(void* arg) {
    (void)arg;
    for (;;) {
        const int s = __atomic_fetch_add(&g_par.next, g_par.chunk, __ATOMIC_RELAXED);
        if (s >= g_par.n) break;
        const int e = s + g_par.chunk < g_par.n ? s + g_par.chunk : g_par.n;
        for (int i = s; i < e; i++) g_par.body(i);
    }
    return NULL;
}
int ref_set_threads(int n) {
    if (n <= 0) n = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n > 256) n = 256;
    if (n < 1) n = 1;
    g_threads = n;
    return n;
}
int ref_num_threads(void) { return g_threads ? g_threads : ref_set_threads(0); }
static void par_for(int n, int chunk, ParBody body) {
    const int T = ref_num_threads();
    g_par.body = body; g_par.n = n; g_par.chunk = chunk; g_par.next = 0;
    pthread_t th[256];
    const int nt = T < n ? T : (n > 0 ? n : 1);
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, par_worker, NULL);
    par_worker(NULL);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

/* ---- tier selection ------------------------------------------------------------------------------ */
int ref_set_tier(int avx2) {
    extern void ref_glue_init(void);
    ref_glue_init(); /* everything = *_c */
    if (!avx2) return 0;
    if (!__builtin_cpu_supports("avx2")) return -1;
    svt_sad_loop_kernel                       = svt_sad_loop_kernel_avx2_intrin;
    svt_ext_all_sad_calculation_8x8_16x16     = svt_ext_all_sad_calculation_8x8_16x16_avx2;
    svt_ext_eight_sad_calculation_32x32_64x64 = svt_ext_eight_sad_calculation_32x32_64x64_avx2;
    svt_ext_sad_calculation_8x8_16x16         = svt_ext_sad_calculation_8x8_16x16_avx2_intrin;
    svt_nxm_sad_kernel                        = svt_nxm_sad_kernel_helper_avx2;
    downsample_2d                             = svt_aom_downsample_2d_avx2;
    svt_av1_fwd_txfm2d_8x8                    = svt_av1_fwd_txfm2d_8x8_avx2;
    svt_av1_fwd_txfm2d_16x16                  = svt_av1_fwd_txfm2d_16x16_avx2;
    svt_av1_fwd_txfm2d_32x32                  = svt_av1_fwd_txfm2d_32x32_avx2;
    svt_av1_fwd_txfm2d_64x64                  = svt_av1_fwd_txfm2d_64x64_avx2;
    svt_av1_quantize_fp                       = svt_av1_quantize_fp_avx2;
    svt_av1_quantize_fp_32x32                 = svt_av1_quantize_fp_32x32_avx2;
    svt_av1_quantize_fp_64x64                 = svt_av1_quantize_fp_64x64_avx2;
    svt_av1_quantize_fp_qm                    = svt_av1_quantize_fp_qm_avx2;
    svt_aom_quantize_b                        = svt_aom_quantize_b_avx2;
    svt_av1_quantize_b_qm                     = svt_av1_quantize_b_qm_avx2;
    svt_cdef_filter_block                     = svt_cdef_filter_block_avx2;
    svt_aom_cdef_find_dir                     = svt_aom_cdef_find_dir_avx2;
    svt_aom_cdef_find_dir_dual                = svt_aom_cdef_find_dir_dual_avx2;
    svt_compute_cdef_dist_8bit                = svt_aom_compute_cdef_dist_8bit_avx2;
    svt_compute_cdef_dist_16bit               = svt_aom_compute_cdef_dist_16bit_avx2;
    svt_aom_copy_rect8_8bit_to_16bit          = svt_aom_copy_rect8_8bit_to_16bit_avx2;
    svt_av1_compute_stats                     = svt_av1_compute_stats_avx2;
    svt_av1_wiener_convolve_add_src           = svt_av1_wiener_convolve_add_src_avx2;
    return 1;
}

/* ---- open-loop ME for one picture ---------------------------------------------------------------
 * restates hme_level_0/1/2 (motion_estimation.c:820-1113), set_final_seach_centre_sb (:2182-2390),
 * check_00_center (:1139-1210), integer_search_b64 (:1249-1520), open_loop_me_fullpel_search_sblock
 * (:781-817) for the controls the B200 T2 path honours. */
typedef struct { const uint8_t* plane[3]; int32_t stride[3], org_x[3], org_y[3], width[3], height[3], reserved[2]; } RefMePicture;
typedef struct { int32_t hme_l0_sa_w, hme_l0_sa_h, hme_l1_sa_w, hme_l1_sa_h, hme_l2_sa_w, hme_l2_sa_h, me_sa_w, me_sa_h, hme_sub_sad, me_sub_sad, check_zero_centre, reserved; } RefMeParams;

static void hme_clip(int16_t org, int16_t* origin, int16_t* sa, int16_t pad, int16_t pic, int round8) {
    if ((int16_t)(org + *origin) < -pad) {
        *origin = (int16_t)(-pad - org);
        *sa     = (int16_t)(*sa - (-pad - (org + *origin)));
    }
    if ((int16_t)(org + *origin) > (int16_t)(pic - 1)) *origin = (int16_t)(*origin - ((org + *origin) - (pic - 1)));
    if ((int16_t)(org + *origin + *sa) > pic) {
        int16_t v = (int16_t)(*sa - ((org + *origin + *sa) - pic));
        *sa = v > 1 ? v : 1;
    }
    if (round8) *sa = (*sa < 8) ? *sa : (int16_t)(*sa & ~7);
}
static int z16(int y16, int x16) { return 4 * (2 * (y16 >> 1) + (x16 >> 1)) + 2 * (y16 & 1) + (x16 & 1); }

static void fullpel_b64(const uint8_t* src, uint32_t ss, const uint8_t* ref, uint32_t rs, int sa_w, int sa_h, int org_x, int org_y, int sub,
                        uint32_t* best_sad, uint32_t* best_mv) {
    uint32_t e16[16][8], e8[64][8], e32[4][8], s16[16], s8[64], s32[4];
    for (int i = 0; i < 85; i++) { best_sad[i] = 128 * 128 * 255; best_mv[i] = 0; }
    uint32_t *b64 = best_sad, *b32 = best_sad + 1, *b16 = best_sad + 5, *b8 = best_sad + 21;
    uint32_t *m64 = best_mv, *m32 = best_mv + 1, *m16 = best_mv + 5, *m8 = best_mv + 21;
    const int w8 = sa_w - (sa_w & 7);
    for (int y = 0; y < sa_h; y++) {
        for (int x = 0; x < w8; x += 8) {
            const uint32_t mv = ((uint32_t)((org_y + y) & 0xffff) << 16) | (uint32_t)((org_x + x) & 0xffff);
            svt_ext_all_sad_calculation_8x8_16x16((uint8_t*)src, ss, (uint8_t*)ref + (size_t)y * rs + x, rs, mv, b8, b16, m8, m16, e16, e8, sub);
            svt_ext_eight_sad_calculation_32x32_64x64(e16, b32, b64, m32, m64, mv, e32);
        }
        for (int x = w8; x < sa_w; x++) {
            const uint32_t mv = ((uint32_t)((org_y + y) & 0xffff) << 16) | (uint32_t)((org_x + x) & 0xffff);
            for (int blk = 0; blk < 16; blk++) {
                const int y16 = blk >> 2, x16 = blk & 3, i16 = z16(y16, x16);
                svt_ext_sad_calculation_8x8_16x16((uint8_t*)src + 16 * y16 * ss + 16 * x16, ss,
                                                  (uint8_t*)ref + (size_t)(y + 16 * y16) * rs + x + 16 * x16, rs, b8 + 4 * i16, b16 + i16,
                                                  m8 + 4 * i16, m16 + i16, mv, &s16[i16], &s8[4 * i16], sub);
            }
            svt_ext_sad_calculation_32x32_64x64(s16, b32, b64, m32, m64, mv, s32);
        }
    }
}

static struct { const RefMePicture* cur; const RefMePicture* refs; const RefMeParams* prm; int n_refs; uint32_t* best_sad; uint32_t* best_mv; int16_t* hme_centre; uint64_t* hme_sad; } g_ref_me_picture;
static void ref_me_picture_body(int i) {
    const RefMePicture* cur = g_ref_me_picture.cur;
    const RefMePicture* refs = g_ref_me_picture.refs;
    const RefMeParams* prm = g_ref_me_picture.prm;
    int n_refs = g_ref_me_picture.n_refs;
    uint32_t* best_sad = g_ref_me_picture.best_sad;
    uint32_t* best_mv = g_ref_me_picture.best_mv;
    int16_t* hme_centre = g_ref_me_picture.hme_centre;
    uint64_t* hme_sad = g_ref_me_picture.hme_sad;
    const int W = cur->width[2], H = cur->height[2], b64_w = (W + 63) >> 6, b64_h = (H + 63) >> 6, nb = b64_w * b64_h;
    {
        const int r = i / nb, b = i % nb, bx = b % b64_w, by = b / b64_w;
        const RefMePicture* rp = &refs[r];
        const RefMeParams*  p  = &prm[r];
        const int sub = p->hme_sub_sad ? 1 : 0;
        int16_t  px[4] = {0, 0, 0, 0}, py[4] = {0, 0, 0, 0};
        uint64_t ls[4] = {0, 0, 0, 0};
        for (int level = 0; level < 3; level++) {
            const int sh = 2 - level;
            const int16_t org_x = (int16_t)((bx * 64) >> sh), org_y = (int16_t)((by * 64) >> sh);
            const int blk_w = (W - bx * 64 < 64 ? W - bx * 64 : 64) >> sh, blk_h = (H - by * 64 < 64 ? H - by * 64 : 64) >> sh;
            int16_t nx[4], ny[4];
            for (int reg = 0; reg < 4; reg++) {
                const int sr_w = reg & 1, sr_h = reg >> 1;
                int16_t sa_w, sa_h, ox, oy, pw, ph;
                if (level == 0) {
                    sa_w = (int16_t)((p->hme_l0_sa_w + 7) & ~7); sa_h = (int16_t)p->hme_l0_sa_h;
                    ox = (int16_t)(-(int16_t)((sa_w * 2) >> 1) + sa_w * sr_w); oy = (int16_t)(-(int16_t)((sa_h * 2) >> 1) + sa_h * sr_h);
                    pw = (int16_t)(rp->org_x[0] - 1); ph = (int16_t)(rp->org_y[0] - 1);
                } else if (level == 1) {
                    sa_w = (int16_t)((p->hme_l1_sa_w + 7) & ~7); sa_h = (int16_t)p->hme_l1_sa_h;
                    ox = (int16_t)(-(sa_w >> 1) + (px[reg] >> 1)); oy = (int16_t)(-(sa_h >> 1) + (py[reg] >> 1));
                    pw = (int16_t)(rp->org_x[1] - 1); ph = (int16_t)(rp->org_y[1] - 1);
                } else {
                    sa_w = (int16_t)((p->hme_l2_sa_w + 7) & ~7); sa_h = (int16_t)p->hme_l2_sa_h;
                    ox = (int16_t)(-(sa_w >> 1) + px[reg]); oy = (int16_t)(-(sa_h >> 1) + py[reg]);
                    pw = ph = 63;
                }
                hme_clip(org_x, &ox, &sa_w, pw, (int16_t)rp->width[level], 1);
                hme_clip(org_y, &oy, &sa_h, ph, (int16_t)rp->height[level], 0);
                const uint8_t* s = cur->plane[level] + (size_t)(cur->org_y[level] + org_y) * cur->stride[level] + cur->org_x[level] + org_x;
                const uint8_t* q = rp->plane[level] + (size_t)(rp->org_y[level] + org_y + oy) * rp->stride[level] + rp->org_x[level] + org_x + ox;
                uint64_t bs = 0;
                int16_t  x = 0, y = 0;
                svt_sad_loop_kernel((uint8_t*)s, (uint32_t)(cur->stride[level] << sub), (uint8_t*)q, (uint32_t)(rp->stride[level] << sub),
                                    (uint32_t)(blk_h >> sub), (uint32_t)blk_w, &bs, &x, &y, (uint32_t)rp->stride[level], 0, sa_w, sa_h);
                if (sub) bs *= 2;
                const int mul = level == 0 ? 4 : (level == 1 ? 2 : 1);
                nx[reg] = (int16_t)((int16_t)(x + ox) * mul);
                ny[reg] = (int16_t)((int16_t)(y + oy) * mul);
                ls[reg] = bs;
            }
            memcpy(px, nx, sizeof(px));
            memcpy(py, ny, sizeof(py));
        }
        int16_t  cx = px[0], cy = py[0];
        uint64_t cs = ls[0];
        for (int reg = 1; reg < 4; reg++)
            if (ls[reg] < cs) { cs = ls[reg]; cx = px[reg]; cy = py[reg]; }
        hme_centre[2 * i] = cx;
        hme_centre[2 * i + 1] = cy;
        hme_sad[i] = cs;
        const int16_t org_x = (int16_t)(bx * 64), org_y = (int16_t)(by * 64);
        const int blk_w = W - org_x < 64 ? W - org_x : 64, blk_h = H - org_y < 64 ? H - org_y : 64;
        const uint8_t* src = cur->plane[2] + (size_t)(cur->org_y[2] + org_y) * cur->stride[2] + cur->org_x[2] + org_x;
        int16_t sx = cx, sy = cy;
        if (p->check_zero_centre && (sx || sy)) {
            const int16_t RW = (int16_t)rp->width[2], RH = (int16_t)rp->height[2];
            if ((int16_t)(org_x + sx) < -63) sx = (int16_t)(-63 - org_x);
            if ((int16_t)(org_x + sx) > RW - 1) sx = (int16_t)(sx - ((org_x + sx) - (RW - 1)));
            if ((int16_t)(org_y + sy) < -63) sy = (int16_t)(-63 - org_y);
            if ((int16_t)(org_y + sy) > RH - 1) sy = (int16_t)(sy - ((org_y + sy) - (RH - 1)));
            const uint8_t* r0 = rp->plane[2] + (size_t)(rp->org_y[2] + org_y) * rp->stride[2] + rp->org_x[2] + org_x;
            uint32_t z  = svt_nxm_sad_kernel(src, (uint32_t)cur->stride[2] << 1, r0, (uint32_t)rp->stride[2] << 1, (uint32_t)blk_h >> 1, (uint32_t)blk_w) << 1;
            uint32_t hs = svt_nxm_sad_kernel(src, (uint32_t)cur->stride[2] << 1, r0 + (ptrdiff_t)sy * rp->stride[2] + sx, (uint32_t)rp->stride[2] << 1,
                                             (uint32_t)blk_h >> 1, (uint32_t)blk_w) << 1;
            if (z <= hs) sx = sy = 0;
        }
        int16_t sa_w = (int16_t)(((p->me_sa_w > 1 ? p->me_sa_w : 1) + 7) & ~7), sa_h = (int16_t)(p->me_sa_h > 3 ? p->me_sa_h : 3);
        int16_t ox = (int16_t)(sx - (sa_w >> 1)), oy = (int16_t)(sy - (sa_h >> 1));
        if ((int16_t)(org_x + ox) < -63) ox = (int16_t)(-63 - org_x);
        if ((int16_t)(org_x + ox) > W - 1) ox = (int16_t)(ox - ((org_x + ox) - (W - 1)));
        if ((int16_t)(org_x + ox + sa_w) > W) { int v = sa_w - ((org_x + ox + sa_w) - W); sa_w = (int16_t)(v > 1 ? v : 1); }
        sa_w = sa_w < 8 ? sa_w : (int16_t)(sa_w & ~7);
        if ((int16_t)(org_y + oy) < -63) oy = (int16_t)(-63 - org_y);
        if ((int16_t)(org_y + oy) > H - 1) oy = (int16_t)(oy - ((org_y + oy) - (H - 1)));
        if ((int16_t)(org_y + oy + sa_h) > H) { int v = sa_h - ((org_y + oy + sa_h) - H); sa_h = (int16_t)(v > 1 ? v : 1); }
        const uint8_t* q = rp->plane[2] + (ptrdiff_t)(rp->org_y[2] + org_y + oy) * rp->stride[2] + rp->org_x[2] + org_x + ox;
        fullpel_b64(src, (uint32_t)cur->stride[2], q, (uint32_t)rp->stride[2], sa_w, sa_h, ox, oy, p->me_sub_sad ? 1 : 0, best_sad + (size_t)i * 85,
                    best_mv + (size_t)i * 85);
        }
}
void ref_me_picture(const RefMePicture* cur, const RefMePicture* refs, const RefMeParams* prm, int n_refs, uint32_t* best_sad,
                    uint32_t* best_mv, int16_t* hme_centre, uint64_t* hme_sad) {
    g_ref_me_picture.cur = cur;
    g_ref_me_picture.refs = refs;
    g_ref_me_picture.prm = prm;
    g_ref_me_picture.n_refs = n_refs;
    g_ref_me_picture.best_sad = best_sad;
    g_ref_me_picture.best_mv = best_mv;
    g_ref_me_picture.hme_centre = hme_centre;
    g_ref_me_picture.hme_sad = hme_sad;
    const int W = cur->width[2], H = cur->height[2], b64_w = (W + 63) >> 6, b64_h = (H + 63) >> 6, nb = b64_w * b64_h;
    par_for(n_refs * nb, 4, ref_me_picture_body);

}


/* ---- transform / quantize / inverse over a block list ---------------------------------------------
 * the trio of svt_aom_estimate_transform -> svt_aom_quantize_inv_quantize -> inverse,
 * one call per block, 8-bit pixels, "fp" quantizer with quantisation matrices (PSY default). */
typedef struct { uint64_t src_off, dst_off; uint32_t src_stride; uint8_t tx_size, tx_type; uint16_t reserved; } RefFwdItem;
typedef struct { uint64_t coef_off, pred_off, recon_off; uint32_t pred_stride, recon_stride; uint8_t tx_size, tx_type, bd, reserved; uint32_t reserved2; } RefInvItem;
typedef struct {
    uint64_t coeff_off, q_off, dq_off; uint32_t scan_off, qm_off, iqm_off, n_coeffs;
    int16_t zbin[2], round[2], quant[2], quant_shift[2], dequant[2]; uint8_t mode, log_scale; uint16_t reserved;
} RefQuantItem;
static const int TXW[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
static const int TXH[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};

static void fwd_one(int16_t* in, int32_t* out, uint32_t stride, int ty, int sz) {
    switch (sz) {
    case 0: svt_av1_transform_two_d_4x4_c(in, out, stride, ty, 8); break; /* no intrinsics-only AVX2 4x4 */
    case 1: svt_av1_fwd_txfm2d_8x8(in, out, stride, ty, 8); break;
    case 2: svt_av1_fwd_txfm2d_16x16(in, out, stride, ty, 8); break;
    case 3: svt_av1_fwd_txfm2d_32x32(in, out, stride, ty, 8); break;
    case 4: svt_av1_fwd_txfm2d_64x64(in, out, stride, ty, 8); break;
    case 5: svt_av1_fwd_txfm2d_4x8(in, out, stride, ty, 8); break;
    case 6: svt_av1_fwd_txfm2d_8x4(in, out, stride, ty, 8); break;
    case 7: svt_av1_fwd_txfm2d_8x16(in, out, stride, ty, 8); break;
    case 8: svt_av1_fwd_txfm2d_16x8(in, out, stride, ty, 8); break;
    case 9: svt_av1_fwd_txfm2d_16x32(in, out, stride, ty, 8); break;
    case 10: svt_av1_fwd_txfm2d_32x16(in, out, stride, ty, 8); break;
    case 11: svt_av1_fwd_txfm2d_32x64(in, out, stride, ty, 8); break;
    case 12: svt_av1_fwd_txfm2d_64x32(in, out, stride, ty, 8); break;
    case 13: svt_av1_fwd_txfm2d_4x16(in, out, stride, ty, 8); break;
    case 14: svt_av1_fwd_txfm2d_16x4(in, out, stride, ty, 8); break;
    case 15: svt_av1_fwd_txfm2d_8x32(in, out, stride, ty, 8); break;
    case 16: svt_av1_fwd_txfm2d_32x8(in, out, stride, ty, 8); break;
    case 17: svt_av1_fwd_txfm2d_16x64(in, out, stride, ty, 8); break;
    default: svt_av1_fwd_txfm2d_64x16(in, out, stride, ty, 8); break;
    }
}


static struct { const int16_t* residual; int32_t* coeff; const RefFwdItem* items; } g_fwd;
static void fwd_body(int i) {
    const RefFwdItem* it = &g_fwd.items[i];
    const int W = TXW[it->tx_size], H = TXH[it->tx_size];
    if ((it->reserved & 1) && (W > 32 || H > 32)) {
        /* packed output: the re-pack half of svt_handle_transform64x64 & co (transforms.c:2374-2542) */
        DECLARE_ALIGNED(64, int32_t, tmp[64 * 64]);
        const int Wp = W > 32 ? 32 : W, Hp = H > 32 ? 32 : H;
        fwd_one((int16_t*)g_fwd.residual + it->src_off, tmp, it->src_stride, it->tx_type, it->tx_size);
        for (int r = 0; r < Hp; r++) memcpy(g_fwd.coeff + it->dst_off + (size_t)r * Wp, tmp + (size_t)r * W, (size_t)Wp * sizeof(int32_t));
        return;
    }
    fwd_one((int16_t*)g_fwd.residual + it->src_off, g_fwd.coeff + it->dst_off, it->src_stride, it->tx_type, it->tx_size);
}
void ref_fwd_txfm_batch(const int16_t* residual, int32_t* coeff, const RefFwdItem* items, int n) {
    g_fwd.residual = residual; g_fwd.coeff = coeff; g_fwd.items = items;
    par_for(n, 64, fwd_body);
}

static struct { const int32_t* coeff; int32_t *q, *dq; const int16_t *scan, *iscan; const uint8_t* qm; const RefQuantItem* items; uint16_t* eobs; } g_q;
static void quant_body(int i) {
    const int32_t* coeff = g_q.coeff;
    int32_t *q = g_q.q, *dq = g_q.dq;
    uint16_t* eobs = g_q.eobs;
    const RefQuantItem* it = &g_q.items[i];
    const uint8_t* wm = it->qm_off == 0xffffffffu ? NULL : g_q.qm + it->qm_off;
    const uint8_t* im = it->iqm_off == 0xffffffffu ? NULL : g_q.qm + it->iqm_off;
    const int16_t* sc = g_q.scan + it->scan_off;
    const int16_t* isc = g_q.iscan + it->scan_off; /* the SIMD tiers derive eob from the inverse scan */
    /* MacroblockPlane tables are int16[8] = {DC, AC x 7}; the SIMD tiers load all eight lanes */
    DECLARE_ALIGNED(16, int16_t, zbin[8]);
    DECLARE_ALIGNED(16, int16_t, round[8]);
    DECLARE_ALIGNED(16, int16_t, quant[8]);
    DECLARE_ALIGNED(16, int16_t, quant_shift[8]);
    DECLARE_ALIGNED(16, int16_t, dequant[8]);
    for (int k = 0; k < 8; k++) {
        zbin[k] = it->zbin[k != 0]; round[k] = it->round[k != 0]; quant[k] = it->quant[k != 0];
        quant_shift[k] = it->quant_shift[k != 0]; dequant[k] = it->dequant[k != 0];
    }
    if (it->mode == 2) { /* SVT_B200_QUANT_FP_LBD */
        if (wm || im)
            svt_av1_quantize_fp_qm(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                   dq + it->dq_off, dequant, &eobs[i], sc, isc, wm, im, it->log_scale);
        else if (it->log_scale == 0)
            svt_av1_quantize_fp(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                dq + it->dq_off, dequant, &eobs[i], sc, isc);
        else if (it->log_scale == 1)
            svt_av1_quantize_fp_32x32(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                      dq + it->dq_off, dequant, &eobs[i], sc, isc);
        else
            svt_av1_quantize_fp_64x64(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                      dq + it->dq_off, dequant, &eobs[i], sc, isc);
    } else if (it->mode == 0)
        svt_aom_quantize_b(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off, dq + it->dq_off,
                           dequant, &eobs[i], sc, isc, wm, im, it->log_scale);
    else if (it->mode == 1)
        svt_aom_highbd_quantize_b(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                  dq + it->dq_off, dequant, &eobs[i], sc, isc, wm, im, it->log_scale);
    else
        svt_av1_highbd_quantize_fp_qm(coeff + it->coeff_off, it->n_coeffs, zbin, round, quant, quant_shift, q + it->q_off,
                                      dq + it->dq_off, dequant, &eobs[i], sc, isc, wm, im, it->log_scale);
}
void ref_quant_batch(const int32_t* coeff, int32_t* q, int32_t* dq, const int16_t* scan, const int16_t* iscan, const uint8_t* qm,
                     const RefQuantItem* items, int n, uint16_t* eobs) {
    g_q.coeff = coeff; g_q.q = q; g_q.dq = dq; g_q.scan = scan; g_q.iscan = iscan; g_q.qm = qm; g_q.items = items; g_q.eobs = eobs;
    par_for(n, 64, quant_body);
}

static struct { const int32_t* coeff; const uint8_t* pred; uint8_t* recon; const RefInvItem* items; } g_inv;
static void inv_body(int i) {
    const RefInvItem* it = &g_inv.items[i];
    TxfmParam tp;
    memset(&tp, 0, sizeof(tp));
    tp.tx_type = it->tx_type;
    tp.tx_size = it->tx_size;
    tp.eob = TXW[it->tx_size] * TXH[it->tx_size]; /* full block */
    if (tp.eob > 1024) tp.eob = 1024;
    tp.bd = 8;
    tp.is_hbd = 1;
    svt_av1_inv_txfm_add((const TranLow*)(g_inv.coeff + it->coef_off), (uint8_t*)g_inv.pred + it->pred_off, (int32_t)it->pred_stride,
                         g_inv.recon + it->recon_off, (int32_t)it->recon_stride, &tp);
}
void ref_inv_txfm_batch_8bit(const int32_t* coeff, const uint8_t* pred, uint8_t* recon, const RefInvItem* items, int n) {
    g_inv.coeff = coeff; g_inv.pred = pred; g_inv.recon = recon; g_inv.items = items;
    par_for(n, 64, inv_body);
}

/* ---- CDEF picture search / apply (cdef_seg_search, cdef_process.c:106-352; svt_av1_cdef_frame) ----- */
typedef struct {
    const void *recon_y, *recon_cb, *recon_cr, *src_y, *src_cb, *src_cr;
    int32_t recon_stride_y, recon_stride_c, src_stride_y, src_stride_c, width, height, bit_depth, damping, subsampling_factor, reserved;
} RefCdefFrame;

static void cdef_tile(uint16_t* inbuf, const uint8_t* plane, int stride, int fbr, int fbc, int nvfb, int nhfb, int fbs, int vsz, int hsz) {
    uint16_t* in = inbuf + CDEF_VBORDER * CDEF_BSTRIDE + CDEF_HBORDER;
    for (int i = 0; i < CDEF_BSTRIDE * (64 + 2 * CDEF_VBORDER); i++) inbuf[i] = CDEF_VERY_LARGE;
    const int yoff = CDEF_VBORDER * (fbr != 0), xoff = CDEF_HBORDER * (fbc != 0);
    const int ysize = vsz + CDEF_VBORDER * (fbr + 1 < nvfb) + yoff, xsize = hsz + CDEF_HBORDER * (fbc + 1 < nhfb) + xoff;
    svt_aom_copy_rect8_8bit_to_16bit(&in[-yoff * CDEF_BSTRIDE - xoff], CDEF_BSTRIDE, plane + (ptrdiff_t)(fbr * fbs - yoff) * stride + fbc * fbs - xoff,
                                     stride, ysize, xsize);
}
static int cdef_list(const RefCdefFrame* f, const uint8_t* skip8x8, int fbr, int fbc, CdefList* dlist) {
    const int w8 = (f->width + 7) >> 3, h8 = (f->height + 7) >> 3;
    int cnt = 0;
    for (int by = 0; by < 8; by++)
        for (int bx = 0; bx < 8; bx++) {
            const int gy = fbr * 8 + by, gx = fbc * 8 + bx;
            if (gy < h8 && gx < w8 && !skip8x8[gy * w8 + gx]) { dlist[cnt].by = (uint8_t)by; dlist[cnt].bx = (uint8_t)bx; cnt++; }
        }
    return cnt;
}

static struct { const RefCdefFrame* f; const uint8_t* skip; const int *sy, *su; int ng; uint64_t* mse; uint8_t* dir; int32_t* var; } g_cs;
static void cdef_search_body(int fb) {
    const RefCdefFrame* f = g_cs.f;
    const int ng = g_cs.ng;
    uint64_t* mse = g_cs.mse;
    const int nhfb = (f->width + 63) >> 6, nvfb = (f->height + 63) >> 6, nfb = nhfb * nvfb;
    const int fbr = fb / nhfb, fbc = fb % nhfb;
    CdefList dlist[64];
    const int cnt = cdef_list(f, g_cs.skip, fbr, fbc, dlist);
    if (!cnt) {
        for (int g = 0; g < ng; g++) mse[(size_t)fb * ng + g] = mse[(size_t)(nfb + fb) * ng + g] = 0;
        return;
    }
    DECLARE_ALIGNED(32, uint16_t, inbuf[CDEF_INBUF_SIZE]);
    DECLARE_ALIGNED(32, uint16_t, tmp_dst[1 << (MAX_SB_SIZE_LOG2 * 2)]);
    uint8_t dir[CDEF_NBLOCKS][CDEF_NBLOCKS];
    int32_t var[CDEF_NBLOCKS][CDEF_NBLOCKS];
    int32_t dirinit = 0;
    for (int pli = 0; pli < 3; pli++) {
        const int dec = pli ? 1 : 0, fbs = 64 >> dec, pw = f->width >> dec, ph = f->height >> dec;
        const uint8_t* rec = (const uint8_t*)(pli == 0 ? f->recon_y : (pli == 1 ? f->recon_cb : f->recon_cr));
        const uint8_t* src = (const uint8_t*)(pli == 0 ? f->src_y : (pli == 1 ? f->src_cb : f->src_cr));
        const int rs = pli ? f->recon_stride_c : f->recon_stride_y, ss = pli ? f->src_stride_c : f->src_stride_y;
        const int hsz = fbs < pw - fbc * fbs ? fbs : pw - fbc * fbs, vsz = fbs < ph - fbr * fbs ? fbs : ph - fbr * fbs;
        cdef_tile(inbuf, rec, rs, fbr, fbc, nvfb, nhfb, fbs, vsz, hsz);
        uint16_t* in = inbuf + CDEF_VBORDER * CDEF_BSTRIDE + CDEF_HBORDER;
        int subs = f->subsampling_factor;
        if (subs > (dec ? 1 : 4)) subs = dec ? 1 : 4;
        for (int g = 0; g < ng; g++) {
            uint64_t* m = &mse[(size_t)((pli ? 1 : 0) * nfb + fb) * ng + g];
            const int sv = pli ? g_cs.su[g] : g_cs.sy[g];
            if (sv < 0) { *m = (uint64_t)1040400 * 64; continue; }
            const int pri = sv / CDEF_SEC_STRENGTHS, sec = sv % CDEF_SEC_STRENGTHS;
            svt_cdef_filter_fb((uint8_t*)tmp_dst, NULL, 0, in, dec, dec, dir, &dirinit, var, pli, dlist, cnt, pri, sec + (sec == 3), f->damping,
                               f->damping, 0, (uint8_t)subs);
            const uint64_t d = svt_compute_cdef_dist_8bit(src + (size_t)(fbr * fbs) * ss + fbc * fbs, ss, (uint8_t*)tmp_dst, dlist, cnt,
                                                          dec ? BLOCK_4X4 : BLOCK_8X8, 0, pli, (uint8_t)subs);
            if (pli == 2) *m += d * subs;
            else *m = d * subs;
        }
    }
    for (int k = 0; k < cnt; k++) {
        g_cs.dir[(size_t)fb * 64 + dlist[k].by * 8 + dlist[k].bx] = dir[dlist[k].by][dlist[k].bx];
        g_cs.var[(size_t)fb * 64 + dlist[k].by * 8 + dlist[k].bx] = var[dlist[k].by][dlist[k].bx];
    }
}
void ref_cdef_search_frame(const RefCdefFrame* f, const uint8_t* skip8x8, const int* str_y, const int* str_uv, int ng, uint64_t* mse, uint8_t* dir_out,
                           int32_t* var_out) {
    g_cs.f = f; g_cs.skip = skip8x8; g_cs.sy = str_y; g_cs.su = str_uv; g_cs.ng = ng; g_cs.mse = mse; g_cs.dir = dir_out; g_cs.var = var_out;
    par_for(((f->width + 63) >> 6) * ((f->height + 63) >> 6), 1, cdef_search_body);
}

static struct { const RefCdefFrame* f; const uint8_t* skip; const int8_t* idx; const int *ys, *us; uint8_t *oy, *ocb, *ocr; int os_y, os_c; } g_ca;
static void cdef_apply_body(int fb) {
    const RefCdefFrame* f = g_ca.f;
    const int nhfb = (f->width + 63) >> 6, nvfb = (f->height + 63) >> 6;
    const int fbr = fb / nhfb, fbc = fb % nhfb;
    if (g_ca.idx[fb] < 0) return;
    const int ys = g_ca.ys[g_ca.idx[fb]], us = g_ca.us[g_ca.idx[fb]];
    if (!ys && !us) return;
    CdefList dlist[64];
    const int cnt = cdef_list(f, g_ca.skip, fbr, fbc, dlist);
    if (!cnt) return;
    DECLARE_ALIGNED(32, uint16_t, inbuf[CDEF_INBUF_SIZE]);
    uint8_t dir[CDEF_NBLOCKS][CDEF_NBLOCKS];
    int32_t var[CDEF_NBLOCKS][CDEF_NBLOCKS];
    int32_t dirinit = 0;
    for (int pli = 0; pli < 3; pli++) {
        const int dec = pli ? 1 : 0, fbs = 64 >> dec, pw = f->width >> dec, ph = f->height >> dec;
        const uint8_t* rec = (const uint8_t*)(pli == 0 ? f->recon_y : (pli == 1 ? f->recon_cb : f->recon_cr));
        uint8_t* out = pli == 0 ? g_ca.oy : (pli == 1 ? g_ca.ocb : g_ca.ocr);
        const int rs = pli ? f->recon_stride_c : f->recon_stride_y, os = pli ? g_ca.os_c : g_ca.os_y;
        const int hsz = fbs < pw - fbc * fbs ? fbs : pw - fbc * fbs, vsz = fbs < ph - fbr * fbs ? fbs : ph - fbr * fbs;
        cdef_tile(inbuf, rec, rs, fbr, fbc, nvfb, nhfb, fbs, vsz, hsz);
        uint16_t* in = inbuf + CDEF_VBORDER * CDEF_BSTRIDE + CDEF_HBORDER;
        const int sv = pli ? us : ys, pri = sv / CDEF_SEC_STRENGTHS, sec = sv % CDEF_SEC_STRENGTHS;
        if (pli == 0 || pri || sec)
            svt_cdef_filter_fb(out + (size_t)(fbr * fbs) * os + fbc * fbs, NULL, os, in, dec, dec, dir, &dirinit, var, pli, dlist, cnt, pri,
                               sec + (sec == 3), f->damping, f->damping, 0, 1);
    }
}
void ref_cdef_apply_frame(const RefCdefFrame* f, const uint8_t* skip8x8, const int8_t* fb_idx, const int* y_str, const int* uv_str, uint8_t* out_y,
                          uint8_t* out_cb, uint8_t* out_cr, int os_y, int os_c) {
    g_ca.f = f; g_ca.skip = skip8x8; g_ca.idx = fb_idx; g_ca.ys = y_str; g_ca.us = uv_str; g_ca.oy = out_y; g_ca.ocb = out_cb; g_ca.ocr = out_cr;
    g_ca.os_y = os_y; g_ca.os_c = os_c;
    par_for(((f->width + 63) >> 6) * ((f->height + 63) >> 6), 1, cdef_apply_body);
}

/* ---- Wiener statistics / filter over unit lists ------------------------------------------------------ */
typedef struct { uint64_t dgd_off, src_off; int32_t dgd_stride, src_stride, h_start, h_end, v_start, v_end, wiener_win, reserved; } RefStatsItem;
typedef struct { uint64_t src_off, dst_off; int32_t src_stride, dst_stride; uint16_t w, h; uint32_t reserved; int16_t hfilter[8], vfilter[8]; } RefWienerUnit;

static struct { const uint8_t *dgd, *src; const RefStatsItem* items; int64_t *M, *H; } g_st;
static void stats_body(int i) {
    int64_t m[WIENER_WIN2], h[WIENER_WIN2 * WIENER_WIN2];
    const RefStatsItem* it = &g_st.items[i];
    svt_av1_compute_stats(it->wiener_win, g_st.dgd + it->dgd_off, g_st.src + it->src_off, it->h_start, it->h_end, it->v_start, it->v_end,
                          it->dgd_stride, it->src_stride, m, h);
    const int w2 = it->wiener_win * it->wiener_win;
    memcpy(g_st.M + (size_t)i * 49, m, sizeof(int64_t) * w2);
    memcpy(g_st.H + (size_t)i * 2401, h, sizeof(int64_t) * w2 * w2);
}
void ref_compute_stats_batch(const uint8_t* dgd, const uint8_t* src, const RefStatsItem* items, int n, int64_t* M, int64_t* H) {
    g_st.dgd = dgd; g_st.src = src; g_st.items = items; g_st.M = M; g_st.H = H;
    par_for(n, 1, stats_body);
}

static struct { const uint8_t* src; uint8_t* dst; const RefWienerUnit* units; } g_wu;
static void wiener_body(int i) {
    const RefWienerUnit* u = &g_wu.units[i];
    /* the reference derives the kernel base by masking the low address bits (convolve.c:48-56) */
    DECLARE_ALIGNED(256, int16_t, fx[128]);
    DECLARE_ALIGNED(256, int16_t, fy[128]);
    memcpy(fx, u->hfilter, 16);
    memcpy(fy, u->vfilter, 16);
    ConvolveParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.round_0 = WIENER_ROUND0_BITS;
    cp.round_1 = 2 * FILTER_BITS - cp.round_0;
    svt_av1_wiener_convolve_add_src(g_wu.src + u->src_off, u->src_stride, g_wu.dst + u->dst_off, u->dst_stride, fx, fy, u->w, u->h, &cp);
}
void ref_wiener_units_8bit(const uint8_t* src, uint8_t* dst, const RefWienerUnit* units, int n) {
    g_wu.src = src; g_wu.dst = dst; g_wu.units = units;
    par_for(n, 16, wiener_body);
}

/* ---- a13 groundwork: one restoration unit through the reference's own stripe loop ------------------------
 * svt_av1_loop_restoration_filter_unit (restoration.c:1067-1135) with need_boundaries = 1, RESTORE_WIENER, 8 bit.
 * `data` / `dst` point at pixel (0,0) of the plane; limits = {h_start, h_end, v_start, v_end}; tile = {left, top,
 * right, bottom}; above/below = the saved stripe-boundary lines (RESTORATION_CTX_VERT rows per stripe, `bstride`
 * bytes per row, logical column x at byte x + RESTORATION_EXTRA_HORZ is handled by the caller's pointer). */
static void lr_filter_unit_8bit(const RestorationUnitInfo* ruip, uint8_t* data, int stride, uint8_t* dst, int dst_stride,
                                const int32_t* limits, uint8_t* above, uint8_t* below, int bstride, const int32_t* tile,
                                int tile_stripe0, int ss_x, int ss_y, int optimized_lr);

void ref_lr_filter_unit_wiener_8bit(uint8_t* data, int stride, uint8_t* dst, int dst_stride, const int32_t* limits,
                                    const int16_t* hfilter, const int16_t* vfilter, uint8_t* above, uint8_t* below, int bstride,
                                    const int32_t* tile, int tile_stripe0, int ss_x, int ss_y, int optimized_lr) {
    RestorationUnitInfo rui;
    memset(&rui, 0, sizeof(rui));
    rui.restoration_type = RESTORE_WIENER;
    memcpy(rui.wiener_info.hfilter, hfilter, 8 * sizeof(int16_t));
    memcpy(rui.wiener_info.vfilter, vfilter, 8 * sizeof(int16_t));
    lr_filter_unit_8bit(&rui, data, stride, dst, dst_stride, limits, above, below, bstride, tile, tile_stripe0, ss_x, ss_y, optimized_lr);
}

/* the same with the self-guided filter (svt_aom_sgrproj_filter_stripe, restoration.c:994-1015) */
void ref_lr_filter_unit_sgrproj_8bit(uint8_t* data, int stride, uint8_t* dst, int dst_stride, const int32_t* limits, int ep,
                                     const int32_t* xqd, uint8_t* above, uint8_t* below, int bstride, const int32_t* tile,
                                     int tile_stripe0, int ss_x, int ss_y, int optimized_lr) {
    RestorationUnitInfo rui;
    memset(&rui, 0, sizeof(rui));
    rui.restoration_type    = RESTORE_SGRPROJ;
    rui.sgrproj_info.ep     = ep;
    rui.sgrproj_info.xqd[0] = xqd[0];
    rui.sgrproj_info.xqd[1] = xqd[1];
    lr_filter_unit_8bit(&rui, data, stride, dst, dst_stride, limits, above, below, bstride, tile, tile_stripe0, ss_x, ss_y, optimized_lr);
}

static void lr_filter_unit_8bit(const RestorationUnitInfo* ruip, uint8_t* data, int stride, uint8_t* dst, int dst_stride,
                                const int32_t* limits, uint8_t* above, uint8_t* below, int bstride, const int32_t* tile,
                                int tile_stripe0, int ss_x, int ss_y, int optimized_lr) {
    RestorationTileLimits lim = {limits[0], limits[1], limits[2], limits[3]};
    RestorationUnitInfo   rui = *ruip;
    RestorationStripeBoundaries rsb;
    rsb.stripe_boundary_above  = above;
    rsb.stripe_boundary_below  = below;
    rsb.stripe_boundary_stride = bstride;
    rsb.stripe_boundary_size   = 0;
    RestorationLineBuffers* rlbs = (RestorationLineBuffers*)malloc(sizeof(RestorationLineBuffers));
    Av1PixelRect            tr   = {tile[0], tile[1], tile[2], tile[3]};
    int32_t*                tmp  = (int32_t*)malloc(RESTORATION_TMPBUF_SIZE);
    svt_av1_loop_restoration_filter_unit(1, &lim, &rui, &rsb, rlbs, &tr, tile_stripe0, ss_x, ss_y, 0, 8, data, stride, dst,
                                         dst_stride, tmp, optimized_lr);
    free(tmp);
    free(rlbs);
}
