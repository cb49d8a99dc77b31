/* oracle/enc_counters.c -- TEST / MEASUREMENT INFRASTRUCTURE.  Call counters around the reference encoder's three transform
 * entry points, linked into oracle/_ref/libsvtav1_enc.so with `-Wl,--wrap=` (the reference objects are untouched: the linker
 * routes their cross-object calls through the __wrap_ functions below, which count and forward to the real ones).
 * tools/measure_tx_search.py reads them after an encode to obtain SURVEY.md 8(d)'s "MD search factor k":
 *     k = sum of N over svt_aom_estimate_transform calls / (1.5 * W * H)   per frame
 * i.e. how many times mode decision pushes each picture sample through forward transform + quantisation.
 *   svt_aom_estimate_transform        Source/Lib/Codec/transforms.c:3158
 *   svt_aom_inv_transform_recon8bit   Source/Lib/Codec/inv_transforms.c:3087
 *   svt_aom_inv_transform_recon       Source/Lib/Codec/inv_transforms.c:3148 */
#include <stdatomic.h>
#include <stdint.h>
#include <string.h>
#include "definitions.h"
#include "transforms.h"
#include "inv_transforms.h"

enum { C_FWD_CALLS, C_FWD_COEFFS, C_FWD_COEFFS_SHAPED, C_INV_CALLS, C_INV_COEFFS, C_INV_EOB0_CALLS, C_FWD_BY_SIZE, C_N = C_FWD_BY_SIZE + TX_SIZES_ALL };
static _Atomic uint64_t g_cnt[C_N];

void ref_counters_reset(void) {
    for (int i = 0; i < C_N; i++) atomic_store(&g_cnt[i], 0);
}
/* out[0..5] as the enum above, out[6..6+19) = forward calls per TxSize */
int ref_counters_read(uint64_t* out, int cap) {
    for (int i = 0; i < C_N && i < cap; i++) out[i] = atomic_load(&g_cnt[i]);
    return C_N;
}

static uint64_t n_of(TxSize s) { return (uint64_t)tx_size_wide[s] * tx_size_high[s]; }

EbErrorType __real_svt_aom_estimate_transform(PictureControlSet* pcs, ModeDecisionContext* ctx, int16_t* residual_buffer, uint32_t residual_stride,
                                              int32_t* coeff_buffer, uint32_t coeff_stride, TxSize transform_size, uint64_t* three_quad_energy,
                                              uint32_t bit_depth, TxType transform_type, PlaneType component_type, EB_TRANS_COEFF_SHAPE shape);
EbErrorType __wrap_svt_aom_estimate_transform(PictureControlSet* pcs, ModeDecisionContext* ctx, int16_t* residual_buffer, uint32_t residual_stride,
                                              int32_t* coeff_buffer, uint32_t coeff_stride, TxSize transform_size, uint64_t* three_quad_energy,
                                              uint32_t bit_depth, TxType transform_type, PlaneType component_type, EB_TRANS_COEFF_SHAPE shape) {
    const uint64_t n = n_of(transform_size);
    atomic_fetch_add(&g_cnt[C_FWD_CALLS], 1);
    atomic_fetch_add(&g_cnt[C_FWD_COEFFS], n);
    /* N2 / N4 shapes compute a quarter / a sixteenth of the outputs (the _N2 / _N4 kernels) */
    atomic_fetch_add(&g_cnt[C_FWD_COEFFS_SHAPED], shape == N2_SHAPE ? n / 4 : shape == N4_SHAPE ? n / 16 : shape == ONLY_DC_SHAPE ? 1 : n);
    atomic_fetch_add(&g_cnt[C_FWD_BY_SIZE + transform_size], 1);
    return __real_svt_aom_estimate_transform(pcs, ctx, residual_buffer, residual_stride, coeff_buffer, coeff_stride, transform_size,
                                             three_quad_energy, bit_depth, transform_type, component_type, shape);
}

EbErrorType __real_svt_aom_inv_transform_recon8bit(int32_t* coeff_buffer, uint8_t* recon_buffer_r, uint32_t recon_stride_r, uint8_t* recon_buffer_w,
                                                   uint32_t recon_stride_w, TxSize txsize, TxType transform_type, PlaneType component_type,
                                                   uint32_t eob, uint8_t lossless);
EbErrorType __wrap_svt_aom_inv_transform_recon8bit(int32_t* coeff_buffer, uint8_t* recon_buffer_r, uint32_t recon_stride_r, uint8_t* recon_buffer_w,
                                                   uint32_t recon_stride_w, TxSize txsize, TxType transform_type, PlaneType component_type,
                                                   uint32_t eob, uint8_t lossless) {
    atomic_fetch_add(&g_cnt[C_INV_CALLS], 1);
    atomic_fetch_add(&g_cnt[C_INV_COEFFS], n_of(txsize));
    if (!eob) atomic_fetch_add(&g_cnt[C_INV_EOB0_CALLS], 1);
    return __real_svt_aom_inv_transform_recon8bit(coeff_buffer, recon_buffer_r, recon_stride_r, recon_buffer_w, recon_stride_w, txsize,
                                                  transform_type, component_type, eob, lossless);
}

EbErrorType __real_svt_aom_inv_transform_recon(int32_t* coeff_buffer, uint8_t* recon_buffer_r, uint32_t recon_stride_r, uint8_t* recon_buffer_w,
                                               uint32_t recon_stride_w, TxSize txsize, uint32_t bit_increment, TxType transform_type,
                                               PlaneType component_type, uint32_t eob, uint8_t lossless);
EbErrorType __wrap_svt_aom_inv_transform_recon(int32_t* coeff_buffer, uint8_t* recon_buffer_r, uint32_t recon_stride_r, uint8_t* recon_buffer_w,
                                               uint32_t recon_stride_w, TxSize txsize, uint32_t bit_increment, TxType transform_type,
                                               PlaneType component_type, uint32_t eob, uint8_t lossless) {
    atomic_fetch_add(&g_cnt[C_INV_CALLS], 1);
    atomic_fetch_add(&g_cnt[C_INV_COEFFS], n_of(txsize));
    if (!eob) atomic_fetch_add(&g_cnt[C_INV_EOB0_CALLS], 1);
    return __real_svt_aom_inv_transform_recon(coeff_buffer, recon_buffer_r, recon_stride_r, recon_buffer_w, recon_stride_w, txsize, bit_increment,
                                              transform_type, component_type, eob, lossless);
}
