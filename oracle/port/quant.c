/* oracle/port/quant.c -- TEST INFRASTRUCTURE: CPU restatement of the reference quantizers, keeping
 * the reference's control flow (end-of-block pre-scan, index list) so that the product's
 * element-wise formulation is checked against the sequential one.  Never linked into the product.
 *
 * Follows: svt_aom_quantize_b_c_ii (Source/Lib/Codec/full_loop.c:29-79), svt_aom_highbd_quantize_b_c
 * (:149-198), quantize_fp_helper_c (:282-342), highbd_quantize_fp_helper_c (:387-453). */
#include "port.h"

#define QM_BITS 5
static int32_t rp2(int32_t v, int n) { return n ? (v + (1 << (n - 1))) >> n : v; }
static int32_t clamp16(int64_t v) { return (int32_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

void port_quantize_b_lbd(const int32_t* c, intptr_t n, const int16_t* zbin, const int16_t* rnd, const int16_t* quant,
                         const int16_t* qshift, int32_t* q, int32_t* dq, const int16_t* deq, uint16_t* eob_ptr,
                         const int16_t* scan, const uint8_t* qm, const uint8_t* iqm, int ls) {
    const int32_t zb[2] = {rp2(zbin[0], ls), rp2(zbin[1], ls)};
    intptr_t      nz = n, eob = -1;
    memset(q, 0, n * sizeof(*q));
    memset(dq, 0, n * sizeof(*dq));
    for (intptr_t i = n - 1; i >= 0; i--) { /* drop the in-dead-zone tail */
        const int rc = scan[i];
        const int32_t v = c[rc] * (qm ? qm[rc] : 32);
        if (v < zb[rc != 0] * 32 && v > -zb[rc != 0] * 32) nz--;
        else break;
    }
    for (intptr_t i = 0; i < nz; i++) {
        const int rc = scan[i], ac = rc != 0;
        const int32_t s = c[rc] < 0 ? -1 : 0, a = (c[rc] ^ s) - s, wt = qm ? qm[rc] : 32;
        if (a * wt >= (zb[ac] << QM_BITS)) {
            int64_t t = clamp16((int64_t)a + rp2(rnd[ac], ls));
            t *= wt;
            const int32_t lvl = (int32_t)(((((t * quant[ac]) >> 16) + t) * qshift[ac]) >> (16 - ls + QM_BITS));
            q[rc] = (lvl ^ s) - s;
            const int32_t d = (deq[ac] * (iqm ? iqm[rc] : 32) + 16) >> QM_BITS;
            const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)d) >> ls;
            dq[rc] = (adq ^ s) - s;
            if (lvl) eob = i;
        }
    }
    *eob_ptr = (uint16_t)(eob + 1);
}

void port_quantize_b_hbd(const int32_t* c, intptr_t n, const int16_t* zbin, const int16_t* rnd, const int16_t* quant,
                         const int16_t* qshift, int32_t* q, int32_t* dq, const int16_t* deq, uint16_t* eob_ptr,
                         const int16_t* scan, const uint8_t* qm, const uint8_t* iqm, int ls) {
    const int32_t zb[2] = {rp2(zbin[0], ls), rp2(zbin[1], ls)};
    static __thread intptr_t keep[4096];
    int nk = 0;
    intptr_t eob = -1;
    memset(q, 0, n * sizeof(*q));
    memset(dq, 0, n * sizeof(*dq));
    for (intptr_t i = 0; i < n; i++) {
        const int rc = scan[i];
        const int32_t v = c[rc] * (qm ? qm[rc] : 32);
        if (v >= zb[rc != 0] * 32 || v <= -zb[rc != 0] * 32) keep[nk++] = i;
    }
    for (int k = 0; k < nk; k++) {
        const int rc = scan[keep[k]], ac = rc != 0;
        const int32_t s = c[rc] < 0 ? -1 : 0, a = (c[rc] ^ s) - s;
        const int64_t tw = ((int64_t)a + rp2(rnd[ac], ls)) * (qm ? qm[rc] : 32);
        const int64_t t2 = ((tw * quant[ac]) >> 16) + tw;
        const int32_t lvl = (int32_t)((t2 * qshift[ac]) >> (16 - ls + QM_BITS));
        q[rc] = (lvl ^ s) - s;
        const int32_t d = (deq[ac] * (iqm ? iqm[rc] : 32) + 16) >> QM_BITS;
        const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)d) >> ls;
        dq[rc] = (adq ^ s) - s;
        if (lvl) eob = keep[k];
    }
    *eob_ptr = (uint16_t)(eob + 1);
}

void port_quantize_fp_lbd(const int32_t* c, intptr_t n, const int16_t* rnd, const int16_t* quant, int32_t* q, int32_t* dq,
                          const int16_t* deq, uint16_t* eob_ptr, const int16_t* scan, const uint8_t* qm,
                          const uint8_t* iqm, int ls) {
    const int r2[2] = {rp2(rnd[0], ls), rp2(rnd[1], ls)};
    int eob = -1;
    memset(q, 0, n * sizeof(*q));
    memset(dq, 0, n * sizeof(*dq));
    for (int i = 0; i < n; i++) {
        const int rc = scan[i], ac = rc != 0;
        const int32_t s = c[rc] < 0 ? -1 : 0;
        int64_t a = (c[rc] ^ s) - s;
        int lvl = 0;
        if (!qm && !iqm) {
            if ((a << (1 + ls)) >= deq[ac]) {
                a = clamp16(a + r2[ac]);
                lvl = (int)((a * quant[ac]) >> (16 - ls));
                if (lvl) {
                    q[rc] = (lvl ^ s) - s;
                    const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)(int32_t)deq[ac]) >> ls;
                    dq[rc] = (adq ^ s) - s;
                }
            }
        } else {
            const int wt = qm ? qm[rc] : 32, d = (deq[ac] * (iqm ? iqm[rc] : 32) + 16) >> QM_BITS;
            if (a * wt >= (deq[ac] << (QM_BITS - (1 + ls)))) {
                a = clamp16(a + r2[ac]);
                lvl = (int)((a * wt * quant[ac]) >> (16 - ls + QM_BITS));
                q[rc] = (lvl ^ s) - s;
                const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)d) >> ls;
                dq[rc] = (adq ^ s) - s;
            }
        }
        if (lvl) eob = i;
    }
    *eob_ptr = (uint16_t)(eob + 1);
}

void port_quantize_fp_hbd(const int32_t* c, intptr_t n, const int16_t* rnd, const int16_t* quant, int32_t* q, int32_t* dq,
                          const int16_t* deq, uint16_t* eob_ptr, const int16_t* scan, const uint8_t* qm,
                          const uint8_t* iqm, int ls) {
    int eob = -1;
    for (int i = 0; i < n; i++) {
        const int rc = scan[i], ac = rc != 0;
        const int32_t s = c[rc] < 0 ? -1 : 0;
        const int64_t a = (c[rc] ^ s) - s;
        q[rc] = dq[rc] = 0;
        if (qm || iqm) {
            const int wt = qm ? qm[rc] : 32, d = (deq[ac] * (iqm ? iqm[rc] : 32) + 16) >> QM_BITS;
            if (a * wt >= (deq[ac] << (QM_BITS - (1 + ls)))) {
                const int lvl = (int)(((a + rp2(rnd[ac], ls)) * quant[ac] * wt) >> (16 - ls + QM_BITS));
                q[rc] = (lvl ^ s) - s;
                const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)d) >> ls;
                dq[rc] = (adq ^ s) - s;
                if (lvl) eob = i;
            }
        } else if ((int32_t)((uint32_t)a << (1 + ls)) >= deq[ac]) {
            const int lvl = (int)(((a + rp2(rnd[ac], ls)) * quant[ac]) >> (16 - ls));
            q[rc] = (lvl ^ s) - s;
            const int32_t adq = (int32_t)((uint32_t)lvl * (uint32_t)(int32_t)deq[ac]) >> ls;
            dq[rc] = (adq ^ s) - s;
            if (lvl) eob = i;
        }
    }
    *eob_ptr = (uint16_t)(eob + 1);
}
