// quant_one.cuh -- the per-coefficient quantise/dequantise arithmetic of K7, shared by quant.cu (batch
// and T1 quantizers) and txfm.cu (the fused transform / quantise / inverse-transform kernel).
// Reference: svt_aom_quantize_b_c_ii (Source/Lib/Codec/full_loop.c:29-79), svt_aom_highbd_quantize_b_c
// (:149-198), quantize_fp_helper_c (:282-342), highbd_quantize_fp_helper_c (:387-453).
#pragma once
#include "common.cuh"
#include "../../include/svt_b200.h"

namespace b200 {

constexpr int kQmBits = 5;

__device__ __forceinline__ int32_t round_pow2(int32_t v, int n) { return n == 0 ? v : ((v + (1 << (n - 1))) >> n); }

__device__ __forceinline__ void quant_one(const SvtB200QuantItem& it, int32_t coeff, int rc, const uint8_t* qm,
                                          const uint8_t* iqm, int32_t& q, int32_t& dq) {
    const int     ac   = rc != 0;
    const int     ls   = it.log_scale;
    const int32_t sign = coeff < 0 ? -1 : 0;
    const int32_t absc = (coeff ^ sign) - sign;
    const int32_t wt   = qm ? (int32_t)qm[rc] : (1 << kQmBits);
    const int32_t iwt  = iqm ? (int32_t)iqm[rc] : (1 << kQmBits);
    q = dq = 0;
    switch (it.mode) {
    case SVT_B200_QUANT_B_LBD: {
        const int32_t zbin = round_pow2(it.zbin[ac], ls);
        if ((int32_t)((uint32_t)absc * (uint32_t)wt) >= (zbin << kQmBits)) {
            int32_t t = absc + round_pow2(it.round[ac], ls);
            t         = t < -32768 ? -32768 : (t > 32767 ? 32767 : t);
            long long tmp = (long long)t * wt;
            const int32_t tmp32 = (int32_t)(((((tmp * it.quant[ac]) >> 16) + tmp) * it.quant_shift[ac]) >> (16 - ls + kQmBits));
            q = (tmp32 ^ sign) - sign;
            const int32_t dequant = (it.dequant[ac] * iwt + (1 << (kQmBits - 1))) >> kQmBits;
            const int32_t adq     = (int32_t)((uint32_t)tmp32 * (uint32_t)dequant) >> ls;
            dq = (adq ^ sign) - sign;
        }
        break;
    }
    case SVT_B200_QUANT_B_HBD: {
        const int32_t zbin = round_pow2(it.zbin[ac], ls);
        const int32_t cw   = (int32_t)((uint32_t)coeff * (uint32_t)wt);
        if (cw >= zbin * (1 << kQmBits) || cw <= -zbin * (1 << kQmBits)) {
            const long long tmp1 = (long long)absc + round_pow2(it.round[ac], ls);
            const long long tmpw = tmp1 * wt;
            const long long tmp2 = ((tmpw * it.quant[ac]) >> 16) + tmpw;
            const int32_t   aq   = (int32_t)((tmp2 * it.quant_shift[ac]) >> (16 - ls + kQmBits));
            q = (aq ^ sign) - sign;
            const int32_t dequant = (it.dequant[ac] * iwt + (1 << (kQmBits - 1))) >> kQmBits;
            const int32_t adq     = (int32_t)((uint32_t)aq * (uint32_t)dequant) >> ls;
            dq = (adq ^ sign) - sign;
        }
        break;
    }
    case SVT_B200_QUANT_FP_LBD: {
        const int32_t rnd = round_pow2(it.round[ac], ls);
        if (!qm && !iqm) {
            if (((long long)absc << (1 + ls)) >= (long long)it.dequant[ac]) {
                long long a = (long long)absc + rnd;
                a           = a < -32768 ? -32768 : (a > 32767 ? 32767 : a);
                const int32_t tmp32 = (int32_t)((a * it.quant[ac]) >> (16 - ls));
                if (tmp32) {
                    q = (tmp32 ^ sign) - sign;
                    const int32_t adq = (int32_t)((uint32_t)tmp32 * (uint32_t)(int32_t)it.dequant[ac]) >> ls;
                    dq = (adq ^ sign) - sign;
                }
            }
        } else {
            const int32_t dequant = (it.dequant[ac] * iwt + (1 << (kQmBits - 1))) >> kQmBits;
            if ((long long)absc * wt >= (long long)((int32_t)it.dequant[ac] << (kQmBits - (1 + ls)))) {
                long long a = (long long)absc + rnd;
                a           = a < -32768 ? -32768 : (a > 32767 ? 32767 : a);
                const int32_t tmp32 = (int32_t)((a * wt * it.quant[ac]) >> (16 - ls + kQmBits));
                q = (tmp32 ^ sign) - sign;
                const int32_t adq = (int32_t)((uint32_t)tmp32 * (uint32_t)dequant) >> ls;
                dq = (adq ^ sign) - sign;
            }
        }
        break;
    }
    default: {  // SVT_B200_QUANT_FP_HBD
        const int shift = 16 - ls;
        if (qm || iqm) {
            const int32_t dequant = (it.dequant[ac] * iwt + (1 << (kQmBits - 1))) >> kQmBits;
            if ((long long)absc * wt >= (long long)((int32_t)it.dequant[ac] << (kQmBits - (1 + ls)))) {
                const long long tmp = (long long)absc + round_pow2(it.round[ac], ls);
                const int32_t   aq  = (int32_t)((tmp * it.quant[ac] * wt) >> (shift + kQmBits));
                q = (aq ^ sign) - sign;
                const int32_t adq = (int32_t)((uint32_t)aq * (uint32_t)dequant) >> ls;
                dq = (adq ^ sign) - sign;
            }
        } else {
            if ((int32_t)((uint32_t)absc << (1 + ls)) >= (int32_t)it.dequant[ac]) {
                const long long tmp = (long long)absc + round_pow2(it.round[ac], ls);
                const int32_t   aq  = (int32_t)((tmp * it.quant[ac]) >> shift);
                q = (aq ^ sign) - sign;
                const int32_t adq = (int32_t)((uint32_t)aq * (uint32_t)(int32_t)it.dequant[ac]) >> ls;
                dq = (adq ^ sign) - sign;
            }
        }
        break;
    }
    }
}

}  // namespace b200
