/* oracle/port/sgr.c -- TEST INFRASTRUCTURE: CPU restatement of the self-guided filter and its
 * projection helpers (pixels as uint16 for both bit depths).  Follows Source/Lib/Codec/restoration.c:
 * 669-955 and restoration_pick.c:167-318, 413-498.  The constant tables are the DATA file the product
 * uses (sgr_tables.inc, dumped from the reference).  Never linked into the product. */
#include <math.h>
#include "port.h"
#include "../../svt-av1-psy_b200/csrc/sgr_tables.inc"

static const int PRM[64] = SGR_PARAMS_INIT, XBY[256] = SGR_X_BY_XPLUS1_INIT, OBX[25] = SGR_ONE_BY_X_INIT;
static uint32_t rpu(uint32_t v, int n) { return n ? (v + ((1u << n) >> 1)) >> n : v; }
static int32_t  rps(int32_t v, int n) { return (v + ((1 << n) >> 1)) >> n; }

static void pass(const uint16_t* d, int ds, int w, int h, int r, uint32_t s, int bd, int fast, int32_t* dst, int dstride) {
    static __thread int32_t A[140 * 140], B[140 * 140];
    const int ap = w + 2, n = (2 * r + 1) * (2 * r + 1);
    for (int i = -1; i < h + 1; i += fast ? 2 : 1)
        for (int j = -1; j < w + 1; j++) {
            uint32_t sum = 0, sq = 0;
            for (int dy = -r; dy <= r; dy++)
                for (int dx = -r; dx <= r; dx++) {
                    const uint32_t v = d[(i + dy) * ds + j + dx];
                    sum += v;
                    sq += v * v;
                }
            const uint32_t a = rpu(sq, 2 * (bd - 8)), b = rpu(sum, bd - 8);
            const uint32_t p = (a * n < b * b) ? 0 : a * n - b * b, z = rpu(p * s, 20);
            const int Av = XBY[z < 255 ? z : 255];
            A[(i + 1) * ap + j + 1] = Av;
            B[(i + 1) * ap + j + 1] = (int32_t)rpu((uint32_t)(256 - Av) * sum * (uint32_t)OBX[n - 1], 12);
        }
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int32_t *a0 = A + (i + 1) * ap + j + 1, *b0 = B + (i + 1) * ap + j + 1;
            int32_t a, b, nb = 5;
            if (fast && (i & 1)) {
                nb = 4;
                a = a0[0] * 6 + (a0[-1] + a0[1]) * 5;
                b = b0[0] * 6 + (b0[-1] + b0[1]) * 5;
            } else if (fast) {
                a = (a0[-ap] + a0[ap]) * 6 + (a0[-1 - ap] + a0[-1 + ap] + a0[1 - ap] + a0[1 + ap]) * 5;
                b = (b0[-ap] + b0[ap]) * 6 + (b0[-1 - ap] + b0[-1 + ap] + b0[1 - ap] + b0[1 + ap]) * 5;
            } else {
                a = (a0[0] + a0[-1] + a0[1] + a0[-ap] + a0[ap]) * 4 + (a0[-1 - ap] + a0[-1 + ap] + a0[1 - ap] + a0[1 + ap]) * 3;
                b = (b0[0] + b0[-1] + b0[1] + b0[-ap] + b0[ap]) * 4 + (b0[-1 - ap] + b0[-1 + ap] + b0[1 - ap] + b0[1 + ap]) * 3;
            }
            dst[i * dstride + j] = rps(a * (int32_t)d[i * ds + j] + b, 8 + nb - 4);
        }
}

void port_selfguided(const uint16_t* dgd, int w, int h, int stride, int32_t* flt0, int32_t* flt1, int fs, int idx, int bd) {
    if (PRM[4 * idx] > 0) pass(dgd, stride, w, h, PRM[4 * idx], PRM[4 * idx + 2], bd, 1, flt0, fs);
    if (PRM[4 * idx + 1] > 0) pass(dgd, stride, w, h, PRM[4 * idx + 1], PRM[4 * idx + 3], bd, 0, flt1, fs);
}

void port_sgr_apply(const uint16_t* dat, int w, int h, int stride, int eps, const int32_t* xqd, uint16_t* dst, int ds, int bd) {
    static __thread int32_t f0[128 * 128], f1[128 * 128];
    const int r0 = PRM[4 * eps], r1 = PRM[4 * eps + 1];
    int xq[2];
    if (r0 == 0) { xq[0] = 0; xq[1] = 128 - xqd[1]; }
    else if (r1 == 0) { xq[0] = xqd[0]; xq[1] = 0; }
    else { xq[0] = xqd[0]; xq[1] = 128 - xq[0] - xqd[1]; }
    port_selfguided(dat, w, h, stride, f0, f1, w, eps, bd);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int32_t u = (int32_t)dat[i * stride + j] << 4;
            int32_t v = u << 7;
            if (r0 > 0) v += xq[0] * (f0[i * w + j] - u);
            if (r1 > 0) v += xq[1] * (f1[i * w + j] - u);
            const int16_t o = (int16_t)rps(v, 11);
            dst[i * ds + j] = (uint16_t)(o < 0 ? 0 : (o > (1 << bd) - 1 ? (1 << bd) - 1 : o));
        }
}

int64_t port_pixel_proj_error(const uint16_t* src, int w, int h, int ss, const uint16_t* dat, int ds, const int32_t* f0, int f0s,
                              const int32_t* f1, int f1s, const int32_t* xq, const int32_t* prm, int hbd_form) {
    int64_t err = 0;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int32_t d = dat[i * ds + j], s = src[i * ss + j], u = d << 4;
            int32_t e;
            if (prm[0] > 0 || prm[1] > 0) {
                int32_t v = hbd_form ? (1 << 10) : (u << 7);
                if (prm[0] > 0) v += xq[0] * (f0[i * f0s + j] - u);
                if (prm[1] > 0) v += xq[1] * (f1[i * f1s + j] - u);
                e = hbd_form ? (v >> 11) + d - s : rps(v, 11) - s;
            } else
                e = d - s;
            err += e * e;
        }
    return err;
}

void port_get_proj_subspace(const uint16_t* src, int w, int h, int ss, const uint16_t* dat, int ds, const int32_t* f0, int f0s,
                            const int32_t* f1, int f1s, int* xq, const int32_t* prm) {
    double H[2][2] = {{0, 0}, {0, 0}}, C[2] = {0, 0};
    const int size = w * h;
    xq[0] = xq[1] = 0;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const double u = (double)(dat[i * ds + j] << 4), s = (double)(src[i * ss + j] << 4) - u;
            const double a = prm[0] > 0 ? (double)f0[i * f0s + j] - u : 0, b = prm[1] > 0 ? (double)f1[i * f1s + j] - u : 0;
            H[0][0] += a * a; H[1][1] += b * b; H[0][1] += a * b; C[0] += a * s; C[1] += b * s;
        }
    H[0][0] /= size; H[0][1] /= size; H[1][1] /= size; H[1][0] = H[0][1]; C[0] /= size; C[1] /= size;
    if (prm[0] == 0) {
        if (H[1][1] < 1e-8) return;
        xq[1] = (int)rint(C[1] / H[1][1] * 128);
    } else if (prm[1] == 0) {
        if (H[0][0] < 1e-8) return;
        xq[0] = (int)rint(C[0] / H[0][0] * 128);
    } else {
        const double det = H[0][0] * H[1][1] - H[0][1] * H[1][0];
        if (det < 1e-8) return;
        xq[0] = (int)rint((H[1][1] * C[0] - H[0][1] * C[1]) / det * 128);
        xq[1] = (int)rint((H[0][0] * C[1] - H[1][0] * C[0]) / det * 128);
    }
}
