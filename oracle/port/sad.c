/* oracle/port/sad.c -- TEST INFRASTRUCTURE: CPU restatement of the reference SAD kernels.
 * Never linked into or called by the product (libsvtav1_b200.so); used by tests/, smoke() and the
 * cpu_baseline leg of bench.py only.  Pinned against the unmodified reference objects
 * (oracle/_ref/libsvtav1_ref.so) by tests/test_oracle_pins.py.
 *
 * Follows: Source/Lib/C_DEFAULT/compute_sad_c.c:20-37 (nxm SAD), :58-101 (sad_loop),
 *          Source/Lib/Codec/motion_estimation.c:98-165,171-205,210-427 (8x8..64x64 pyramid). */
#include "port.h"

uint32_t port_nxm_sad(const uint8_t* src, uint32_t src_stride, const uint8_t* ref, uint32_t ref_stride, uint32_t h,
                      uint32_t w) {
    uint32_t sad = 0;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            int d = (int)src[y * src_stride + x] - (int)ref[y * ref_stride + x];
            sad += (uint32_t)(d < 0 ? -d : d);
        }
    return sad;
}

/* compute_sad_c.c:58-101: raster scan, strict '<', best starts at 0xffffff; even search rows are
 * skipped for 16-wide blocks of height <= 16 when skip_search_line is set. */
void port_sad_loop(const uint8_t* src, uint32_t src_stride, const uint8_t* ref, uint32_t ref_stride, uint32_t bh,
                   uint32_t bw, uint64_t* best_sad, int16_t* xc, int16_t* yc, uint32_t ref_step, uint8_t skip,
                   int16_t sa_w, int16_t sa_h) {
    *best_sad = 0xffffff;
    for (int y = 0; y < sa_h; y++) {
        if (bw == 16 && bh <= 16 && skip && (y & 1) == 0) continue;
        for (int x = 0; x < sa_w; x++) {
            uint32_t sad = port_nxm_sad(src, src_stride, ref + (size_t)y * ref_step + x, ref_stride, bh, bw);
            if (sad < *best_sad) {
                *best_sad = sad;
                *xc       = (int16_t)x;
                *yc       = (int16_t)y;
            }
        }
    }
}
