/* oracle/port/lr_unit.c -- TEST INFRASTRUCTURE: CPU restatement of one restoration unit filtered stripe by stripe
 * with the saved boundary lines standing in for the rows above / below each stripe.
 * Follows Source/Lib/Codec/restoration.c: svt_aom_get_stripe_boundary_info (:257-273),
 * svt_aom_setup_processing_stripe_boundary (:289-371), svt_aom_wiener_filter_stripe (:437-459) and the stripe loop of
 * svt_av1_loop_restoration_filter_unit (:1067-1135).  Never linked into the product.
 *
 * Restated without the reference's save / overwrite / restore of the picture rows: every stripe is filtered from a
 * private copy of the rows it reads, into which the boundary lines are substituted. */
#include <stdlib.h>

#include "port.h"

void port_wiener_convolve(const uint16_t* src, ptrdiff_t ss, uint16_t* dst, ptrdiff_t ds, const int16_t* fx, const int16_t* fy, int w,
                          int h, int round0, int round1, int bd, int lbd);

enum { PROC_UNIT = 64, UNIT_OFFSET = 8, BORDER = 3, CTX_VERT = 2, EXTRA_HORZ = 4 };

void port_sgr_apply(const uint16_t* dat, int w, int h, int stride, int eps, const int32_t* xqd, uint16_t* dst, int ds, int bd);

/* rtype 0: Wiener (hfilter, vfilter); 1: self-guided (ep, xqd) */
static void lr_filter_unit_8bit(int rtype, const uint8_t* data, int stride, uint8_t* dst, int dst_stride, const int32_t* limits,
                                const int16_t* hfilter, const int16_t* vfilter, int ep, const int32_t* xqd, const uint8_t* above,
                                const uint8_t* below, int bstride, const int32_t* tile, int tile_stripe0, int ss_x, int ss_y,
                                int optimized_lr) {
    const int h_start = limits[0], h_end = limits[1], v_start = limits[2], v_end = limits[3];
    const int tile_top = tile[1], tile_bottom = tile[3];
    const int unit_w = h_end - h_start, unit_h = v_end - v_start;
    const int full_stripe = PROC_UNIT >> ss_y, runit_off = UNIT_OFFSET >> ss_y, procw = PROC_UNIT >> ss_x;
    /* private window: rows -3 .. h+3 of the stripe, columns h_start-4 .. h_start+wmax+4 where wmax covers the
     * 16-rounded width svt_aom_wiener_filter_stripe hands to the convolution */
    const int wmax = ((unit_w + 15) & ~15), cols = wmax + 2 * EXTRA_HORZ + 8, pitch = cols;
    uint16_t* win = (uint16_t*)malloc((size_t)(full_stripe + 2 * BORDER) * pitch * sizeof(uint16_t));
    uint16_t* out = (uint16_t*)malloc((size_t)full_stripe * 64 * sizeof(uint16_t));
    for (int i = 0; i < unit_h;) {
        const int vs = v_start + i;
        /* svt_aom_get_stripe_boundary_info on the remaining stripes (v_start = vs) */
        const int first_in_tile = (vs == tile_top);
        const int this_h = full_stripe - (first_in_tile ? runit_off : 0);
        const int last_in_tile = (vs + this_h >= tile_bottom);
        const int copy_above = !first_in_tile, copy_below = !last_in_tile;
        const int tile_stripe = (vs - tile_top + runit_off) / full_stripe;
        const int rsb_row = CTX_VERT * (tile_stripe0 + tile_stripe);
        const int nominal = full_stripe - (tile_stripe == 0 ? runit_off : 0);
        const int h = nominal < v_end - vs ? nominal : v_end - vs;
        /* copy the rows the stripe reads */
        for (int r = -BORDER; r < h + BORDER; r++)
            for (int c = 0; c < cols; c++) win[(r + BORDER) * pitch + c] = data[(ptrdiff_t)(vs + r) * stride + h_start - EXTRA_HORZ + c];
        const int line_w = unit_w + 2 * EXTRA_HORZ; /* substituted span: columns h_start-4 .. h_end+4 */
        if (!optimized_lr) {
            if (copy_above)
                for (int r = -BORDER; r < 0; r++) {
                    const int br = rsb_row + (r + CTX_VERT > 0 ? r + CTX_VERT : 0);
                    for (int c = 0; c < line_w; c++) win[(r + BORDER) * pitch + c] = above[(ptrdiff_t)br * bstride + h_start + c];
                }
            if (copy_below)
                for (int r = 0; r < BORDER; r++) {
                    const int br = rsb_row + (r < CTX_VERT - 1 ? r : CTX_VERT - 1);
                    for (int c = 0; c < line_w; c++) win[(h + r + BORDER) * pitch + c] = below[(ptrdiff_t)br * bstride + h_start + c];
                }
        } else { /* optimized_lr: only the outermost context row is replaced, by its inner neighbour */
            if (copy_above)
                for (int c = 0; c < line_w; c++) win[0 * pitch + c] = win[1 * pitch + c];
            if (copy_below)
                for (int c = 0; c < line_w; c++) win[(h + 2 + BORDER) * pitch + c] = win[(h + 1 + BORDER) * pitch + c];
        }
        /* svt_aom_wiener_filter_stripe: 64-column processing units, the last one rounded up to a multiple of 16;
         * svt_aom_sgrproj_filter_stripe: 64-column units, the last one exactly as wide as what is left */
        for (int j = 0; j < unit_w; j += procw) {
            int w = rtype == 0 ? ((unit_w - j + 15) & ~15) : unit_w - j;
            if (w > procw) w = procw;
            if (rtype == 0)
                port_wiener_convolve(win + BORDER * pitch + EXTRA_HORZ + j, pitch, out, 64, hfilter, vfilter, w, h, 3, 11, 8, 1);
            else
                port_sgr_apply(win + BORDER * pitch + EXTRA_HORZ + j, w, h, pitch, ep, xqd, out, 64, 8);
            for (int r = 0; r < h; r++)
                for (int c = 0; c < w; c++) dst[(ptrdiff_t)(vs + r) * dst_stride + h_start + j + c] = (uint8_t)out[r * 64 + c];
        }
        i += h;
    }
    free(win);
    free(out);
}

void port_lr_filter_unit_wiener_8bit(const uint8_t* data, int stride, uint8_t* dst, int dst_stride, const int32_t* limits,
                                     const int16_t* hfilter, const int16_t* vfilter, const uint8_t* above, const uint8_t* below,
                                     int bstride, const int32_t* tile, int tile_stripe0, int ss_x, int ss_y, int optimized_lr) {
    lr_filter_unit_8bit(0, data, stride, dst, dst_stride, limits, hfilter, vfilter, 0, NULL, above, below, bstride, tile, tile_stripe0,
                        ss_x, ss_y, optimized_lr);
}

void port_lr_filter_unit_sgrproj_8bit(const uint8_t* data, int stride, uint8_t* dst, int dst_stride, const int32_t* limits, int ep,
                                      const int32_t* xqd, const uint8_t* above, const uint8_t* below, int bstride, const int32_t* tile,
                                      int tile_stripe0, int ss_x, int ss_y, int optimized_lr) {
    lr_filter_unit_8bit(1, data, stride, dst, dst_stride, limits, NULL, NULL, ep, xqd, above, below, bstride, tile, tile_stripe0, ss_x,
                        ss_y, optimized_lr);
}
