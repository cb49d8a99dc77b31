/* oracle/port/hadamard.c -- TEST INFRASTRUCTURE: CPU restatement of the reference Hadamard/SATD.
 * Follows Source/Lib/C_DEFAULT/picture_operators_c.c:175-330 and common_dsp_rtcd.c:70-77. */
#include "port.h"

static void col(const int16_t* s, ptrdiff_t st, int16_t* o, int n) {
    int16_t a[8], b[8];
    for (int i = 0; i < n; i++) a[i] = s[i * st];
    if (n == 4) {
        b[0] = (int16_t)((a[0] + a[1]) >> 1);
        b[1] = (int16_t)((a[0] - a[1]) >> 1);
        b[2] = (int16_t)((a[2] + a[3]) >> 1);
        b[3] = (int16_t)((a[2] - a[3]) >> 1);
        o[0] = (int16_t)(b[0] + b[2]);
        o[1] = (int16_t)(b[1] + b[3]);
        o[2] = (int16_t)(b[0] - b[2]);
        o[3] = (int16_t)(b[1] - b[3]);
        return;
    }
    for (int i = 0; i < 4; i++) {
        b[2 * i]     = (int16_t)(a[2 * i] + a[2 * i + 1]);
        b[2 * i + 1] = (int16_t)(a[2 * i] - a[2 * i + 1]);
    }
    int16_t c[8];
    for (int h = 0; h < 2; h++) {
        c[4 * h + 0] = (int16_t)(b[4 * h + 0] + b[4 * h + 2]);
        c[4 * h + 1] = (int16_t)(b[4 * h + 1] + b[4 * h + 3]);
        c[4 * h + 2] = (int16_t)(b[4 * h + 0] - b[4 * h + 2]);
        c[4 * h + 3] = (int16_t)(b[4 * h + 1] - b[4 * h + 3]);
    }
    static const int where_sum[4] = {0, 7, 3, 4}, where_dif[4] = {2, 6, 1, 5};
    for (int i = 0; i < 4; i++) {
        o[where_sum[i]] = (int16_t)(c[i] + c[4 + i]);
        o[where_dif[i]] = (int16_t)(c[i] - c[4 + i]);
    }
}

static void had_small(const int16_t* src, ptrdiff_t st, int32_t* coeff, int n) {
    int16_t t1[64], t2[64];
    for (int i = 0; i < n; i++) col(src + i, st, t1 + n * i, n);
    for (int i = 0; i < n; i++) col(t1 + i, n, t2 + n * i, n);
    for (int i = 0; i < n * n; i++) coeff[i] = t2[i];
}

void port_hadamard(const int16_t* src, ptrdiff_t st, int32_t* coeff, int n) {
    if (n <= 8) {
        had_small(src, st, coeff, n);
        return;
    }
    const int h = n / 2, q = h * h, sh = n == 16 ? 1 : 2;
    for (int i = 0; i < 4; i++) port_hadamard(src + (i >> 1) * h * st + (i & 1) * h, st, coeff + i * q, h);
    for (int i = 0; i < q; i++) {
        const int32_t a0 = coeff[i], a1 = coeff[q + i], a2 = coeff[2 * q + i], a3 = coeff[3 * q + i];
        const int32_t b0 = (a0 + a1) >> sh, b1 = (a0 - a1) >> sh, b2 = (a2 + a3) >> sh, b3 = (a2 - a3) >> sh;
        coeff[i] = b0 + b2;
        coeff[q + i] = b1 + b3;
        coeff[2 * q + i] = b0 - b2;
        coeff[3 * q + i] = b1 - b3;
    }
}

int port_satd(const int32_t* coeff, int length) {
    int s = 0;
    for (int i = 0; i < length; i++) s += coeff[i] < 0 ? -coeff[i] : coeff[i];
    return s;
}
