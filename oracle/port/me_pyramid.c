/* oracle/port/me_pyramid.c -- TEST INFRASTRUCTURE: CPU restatement of the 85-PU full-pel search.
 * Follows open_loop_me_fullpel_search_sblock (Source/Lib/Codec/motion_estimation.c:781-817) and the
 * SAD pyramid kernels it drives (:98-427): per search position (raster order) the 64 8x8 SADs
 * (8x4 with doubled pitch, x2, in SUB_SAD mode) are summed to 16x16/32x32/64x64 and each of the 85
 * running (best SAD, MV) pairs is updated on a strict '<'.  Output order: me_context.h:54-75. */
#include "port.h"

static int z16(int y16, int x16) { return 4 * (2 * (y16 >> 1) + (x16 >> 1)) + 2 * (y16 & 1) + (x16 & 1); }

void port_fullpel_search(const uint8_t* src, uint32_t ss, const uint8_t* ref, uint32_t rs, int sa_w, int sa_h, int org_x,
                         int org_y, int sub, uint32_t* best_sad, uint32_t* best_mv) {
    for (int i = 0; i < 85; i++) {
        best_sad[i] = 128 * 128 * 255;
        best_mv[i]  = 0;
    }
    for (int y = 0; y < sa_h; y++)
        for (int x = 0; x < sa_w; x++) {
            uint32_t s8[64], s16[16], s32[4], s64 = 0;
            for (int by = 0; by < 8; by++)
                for (int bx = 0; bx < 8; bx++) {
                    uint32_t acc = 0;
                    for (int r = 0; r < 8; r += sub ? 2 : 1)
                        for (int c = 0; c < 8; c++) {
                            int d = (int)src[(8 * by + r) * ss + 8 * bx + c] - (int)ref[(size_t)(y + 8 * by + r) * rs + x + 8 * bx + c];
                            acc += (uint32_t)(d < 0 ? -d : d);
                        }
                    s8[4 * z16(by >> 1, bx >> 1) + 2 * (by & 1) + (bx & 1)] = sub ? acc << 1 : acc;
                }
            for (int p = 0; p < 16; p++) s16[p] = s8[4 * p] + s8[4 * p + 1] + s8[4 * p + 2] + s8[4 * p + 3];
            for (int q = 0; q < 4; q++) {
                s32[q] = s16[4 * q] + s16[4 * q + 1] + s16[4 * q + 2] + s16[4 * q + 3];
                s64 += s32[q];
            }
            const uint32_t mv = ((uint32_t)((org_y + y) & 0xffff) << 16) | (uint32_t)((org_x + x) & 0xffff);
#define UPD(idx, v) if ((v) < best_sad[idx]) { best_sad[idx] = (v); best_mv[idx] = mv; }
            UPD(0, s64)
            for (int q = 0; q < 4; q++) UPD(1 + q, s32[q])
            for (int p = 0; p < 16; p++) UPD(5 + p, s16[p])
            for (int b = 0; b < 64; b++) UPD(21 + b, s8[b])
#undef UPD
        }
}
