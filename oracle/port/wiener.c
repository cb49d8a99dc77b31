/* oracle/port/wiener.c -- TEST INFRASTRUCTURE: CPU restatement of the Wiener filter and statistics.
 * Pixels are passed as uint16 for both bit depths.  Follows Source/Lib/Codec/convolve.c:57-237 and
 * Source/Lib/Codec/restoration_pick.c:659-745.  Never linked into the product. */
#include "port.h"

static int rp2(int v, int n) { return (v + ((1 << n) >> 1)) >> n; }

/* lbd != 0 reproduces the 8-bit entry's h+6 intermediate rows + zeroed last row */
void port_wiener_convolve(const uint16_t* src, ptrdiff_t ss, uint16_t* dst, ptrdiff_t ds, const int16_t* fx, const int16_t* fy, int w,
                          int h, int round0, int round1, int bd, int lbd) {
    static __thread uint16_t tmp[72 * 64];
    const int limit = (1 << (bd + 1 + 7 - round0)) - 1, rows = lbd ? h + 6 : h + 7;
    for (int r = 0; r < h + 7; r++)
        for (int c = 0; c < w; c++) {
            int v = 0;
            if (r < rows) {
                const uint16_t* p = src + (ptrdiff_t)(r - 3) * ss + c - 3;
                int sum = ((int)p[3] << 7) + (1 << (bd + 6));
                for (int k = 0; k < 8; k++) sum += p[k] * fx[k];
                v = rp2(sum, round0);
                v = v < 0 ? 0 : (v > limit ? limit : v);
            }
            tmp[r * 64 + c] = (uint16_t)v;
        }
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            int sum = ((int)tmp[(r + 3) * 64 + c] << 7) - (1 << (bd + round1 - 1));
            for (int k = 0; k < 8; k++)
                if (r + k < h + 7) sum += tmp[(r + k) * 64 + c] * fy[k];
            int v = rp2(sum, round1);
            dst[r * ds + c] = (uint16_t)(v < 0 ? 0 : (v > (1 << bd) - 1 ? (1 << bd) - 1 : v));
        }
}

void port_compute_stats(int win, const uint16_t* dgd, const uint16_t* src, int h_start, int h_end, int v_start, int v_end, int dstride,
                        int sstride, int64_t* M, int64_t* H, int bd) {
    const int win2 = win * win, half = win >> 1, div = bd == 12 ? 16 : (bd == 10 ? 4 : 1);
    uint64_t sum = 0;
    for (int i = v_start; i < v_end; i++)
        for (int j = h_start; j < h_end; j++) sum += dgd[i * dstride + j];
    const int avg = (int)(sum / (uint64_t)((v_end - v_start) * (h_end - h_start)));
    memset(M, 0, sizeof(*M) * win2);
    memset(H, 0, sizeof(*H) * win2 * win2);
    int32_t y[49];
    for (int i = v_start; i < v_end; i++)
        for (int j = h_start; j < h_end; j++) {
            const int x = (int)src[i * sstride + j] - avg;
            int idx = 0;
            for (int k = -half; k <= half; k++)
                for (int l = -half; l <= half; l++) y[idx++] = (int)dgd[(i + l) * dstride + (j + k)] - avg;
            for (int k = 0; k < win2; k++) {
                M[k] += (int64_t)y[k] * x;
                for (int l = k; l < win2; l++) H[k * win2 + l] += (int64_t)y[k] * y[l];
            }
        }
    for (int k = 0; k < win2; k++) {
        M[k] /= div;
        H[k * win2 + k] /= div;
        for (int l = k + 1; l < win2; l++) {
            H[k * win2 + l] /= div;
            H[l * win2 + k] = H[k * win2 + l];
        }
    }
}
