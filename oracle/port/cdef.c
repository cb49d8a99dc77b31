/* oracle/port/cdef.c -- TEST INFRASTRUCTURE: CPU restatement of the reference CDEF path (8/16-bit
 * pixels held as uint16 here).  Never linked into the product.  Pinned against the unmodified
 * reference objects by tests/test_oracle_pins.py.
 *
 * Follows: constrain / adjust_strength / svt_aom_cdef_find_dir_c / svt_cdef_filter_block_c /
 *          svt_cdef_filter_fb (Source/Lib/Codec/cdef.c:85-430), dist_8xn / mse / compute_cdef_dist
 *          (Source/Lib/Codec/enc_cdef.c:23-233), tile build + strength loop of cdef_seg_search
 *          (Source/Lib/Codec/cdef_process.c:106-352). */
#include <math.h>
#include <stdlib.h>
#include "port.h"

#define VL 0x7f7f
#define BS 144 /* CDEF_BSTRIDE */

static int msb(uint32_t n) { return 31 - __builtin_clz(n); }
static int constrain(int diff, int thr, int damp) {
    if (!thr) return 0;
    int shift = damp - msb(thr);
    if (shift < 0) shift = 0;
    int ad = abs(diff), lim = thr - (ad >> shift);
    if (lim < 0) lim = 0;
    int v = ad < lim ? ad : lim;
    return diff < 0 ? -v : v;
}
static int adjust_strength(int strength, int var) {
    int i = (var >> 6) ? msb(var >> 6) : 0;
    if (i > 12) i = 12;
    return var ? (strength * (4 + i) + 8) >> 4 : 0;
}
static const int DIRS[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                  {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};

int port_cdef_find_dir(const uint16_t* img, int stride, int32_t* var, int cs) {
    int cost[8] = {0}, partial[8][15];
    memset(partial, 0, sizeof(partial));
    static const int dv[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            int x = (img[i * stride + j] >> cs) - 128;
            partial[0][i + j] += x;
            partial[1][i + j / 2] += x;
            partial[2][i] += x;
            partial[3][3 + i - j / 2] += x;
            partial[4][7 + i - j] += x;
            partial[5][3 - i / 2 + j] += x;
            partial[6][j] += x;
            partial[7][i / 2 + j] += x;
        }
    for (int i = 0; i < 8; i++) {
        cost[2] += partial[2][i] * partial[2][i];
        cost[6] += partial[6][i] * partial[6][i];
    }
    cost[2] *= dv[8];
    cost[6] *= dv[8];
    for (int i = 0; i < 7; i++) {
        cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * dv[i + 1];
        cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * dv[i + 1];
    }
    cost[0] += partial[0][7] * partial[0][7] * dv[8];
    cost[4] += partial[4][7] * partial[4][7] * dv[8];
    for (int i = 1; i < 8; i += 2) {
        for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
        cost[i] *= dv[8];
        for (int j = 0; j < 3; j++) cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) * dv[2 * j + 2];
    }
    int bc = 0, bd = 0;
    for (int i = 0; i < 8; i++)
        if (cost[i] > bc) {
            bc = cost[i];
            bd = i;
        }
    *var = (bc - cost[(bd + 4) & 7]) >> 10;
    return bd;
}

static int filter_px(const uint16_t* in, int s, int pri, int sec, int dir, int pd, int sd, int cs) {
    const int odd = (pri >> cs) & 1, pt[2] = {odd ? 3 : 4, odd ? 3 : 2}, st[2] = {2, 1};
    const int x = (int16_t)in[0];
    int16_t sum = 0;
    int mx = x, mn = x;
    for (int k = 0; k < 2; k++) {
        const int o[3] = {DIRS[dir][k][0] * s + DIRS[dir][k][1], DIRS[(dir + 2) & 7][k][0] * s + DIRS[(dir + 2) & 7][k][1],
                          DIRS[(dir + 6) & 7][k][0] * s + DIRS[(dir + 6) & 7][k][1]};
        for (int t = 0; t < 3; t++)
            for (int sg = 0; sg < 2; sg++) {
                const int p = (int16_t)in[sg ? -o[t] : o[t]];
                if (t == 0) sum += (int16_t)(pt[k] * constrain(p - x, pri, pd));
                else sum += (int16_t)(st[k] * constrain(p - x, sec, sd));
                if (p != VL && p > mx) mx = p;
                if (p < mn) mn = p;
            }
    }
    int y = (int16_t)x + ((8 + sum - (sum < 0)) >> 4);
    return (int16_t)(y < mn ? mn : (y > mx ? mx : y));
}

/* bsize: BLOCK_4X4=0, 4X8=1, 8X4=2, 8X8=3; in has pitch BS */
void port_cdef_filter_block(uint16_t* dst, int dstride, const uint16_t* in, int pri, int sec, int dir, int pd, int sd, int bsize,
                            int cs, int subs) {
    const int h = 4 << (bsize == 3 || bsize == 1), w = 4 << (bsize == 3 || bsize == 2);
    for (int i = 0; i < h; i += subs)
        for (int j = 0; j < w; j++) dst[i * dstride + j] = (uint16_t)filter_px(in + i * BS + j, BS, pri, sec, dir, pd, sd, cs);
}

static uint64_t dist8(const uint16_t* a, int as, const uint16_t* b, int bs, int cs, int subs) {
    uint64_t sa = 0, sb = 0, a2 = 0, b2 = 0, ab = 0;
    for (int i = 0; i < 8; i += subs)
        for (int j = 0; j < 8; j++) {
            const uint64_t x = a[i * as + j], y = b[i * bs + j];
            sa += x; sb += y; a2 += x * x; b2 += y * y; ab += x * y;
        }
    const uint64_t va = a2 - ((sa * sa + 32) >> 6), vb = b2 - ((sb * sb + 32) >> 6);
    return (uint64_t)floor(.5 + (b2 + a2 - 2 * ab) * .5 * (va + vb + (400 << 2 * cs)) / (sqrt((20000 << 4 * cs) + va * (double)vb)));
}

/* Whole-picture strength search, 4:2:0.  planes are uint16 [3], strides in pixels.
 * mse out: [2][nfb][ng]; dir/var out: [nfb][64]. */
void port_cdef_search_frame(const uint16_t* const rec[3], const int rstride[3], const uint16_t* const src[3], const int sstride[3],
                            int width, int height, int bit_depth, int damping, int subsampling, const uint8_t* skip8x8,
                            const int* str_y, const int* str_uv, int ng, uint64_t* mse, uint8_t* dir_out, int32_t* var_out) {
    const int nhfb = (width + 63) >> 6, nvfb = (height + 63) >> 6, nfb = nhfb * nvfb, w8 = (width + 7) >> 3, h8 = (height + 7) >> 3;
    const int cs = bit_depth > 8 ? bit_depth - 8 : 0;
    static __thread uint16_t inbuf[BS * (64 + 6)];
    uint16_t* in = inbuf + 3 * BS + 8;
    for (int fbr = 0; fbr < nvfb; fbr++)
        for (int fbc = 0; fbc < nhfb; fbc++) {
            const int fb = fbr * nhfb + fbc;
            uint8_t lst[64][2];
            int cnt = 0;
            for (int by = 0; by < 8; by++)
                for (int bx = 0; bx < 8; bx++) {
                    const int gy = fbr * 8 + by, gx = fbc * 8 + bx;
                    if (gy < h8 && gx < w8 && !skip8x8[gy * w8 + gx]) {
                        lst[cnt][0] = by;
                        lst[cnt][1] = bx;
                        cnt++;
                    }
                }
            if (!cnt) {
                for (int g = 0; g < ng; g++) mse[(size_t)fb * ng + g] = mse[(size_t)(nfb + fb) * ng + g] = 0;
                continue;
            }
            int dir[8][8], var[8][8];
            for (int pli = 0; pli < 3; pli++) {
                const int dec = pli ? 1 : 0, fbs = 64 >> dec, pw = width >> dec, ph = height >> dec, bsz = 8 >> dec;
                const int hsz = fbs < pw - fbc * fbs ? fbs : pw - fbc * fbs, vsz = fbs < ph - fbr * fbs ? fbs : ph - fbr * fbs;
                for (int i = 0; i < BS * 70; i++) inbuf[i] = VL;
                const int yoff = 3 * (fbr != 0), xoff = 8 * (fbc != 0);
                const int ysize = vsz + 3 * (fbr + 1 < nvfb) + yoff, xsize = hsz + 8 * (fbc + 1 < nhfb) + xoff;
                for (int r = 0; r < ysize; r++)
                    for (int c = 0; c < xsize; c++)
                        in[(r - yoff) * BS + c - xoff] = rec[pli][(size_t)(fbr * fbs - yoff + r) * rstride[pli] + fbc * fbs - xoff + c];
                if (pli == 0)
                    for (int k = 0; k < cnt; k++) {
                        const int by = lst[k][0], bx = lst[k][1];
                        dir[by][bx] = port_cdef_find_dir(in + 8 * by * BS + 8 * bx, BS, &var[by][bx], cs);
                        dir_out[(size_t)fb * 64 + by * 8 + bx] = (uint8_t)dir[by][bx];
                        var_out[(size_t)fb * 64 + by * 8 + bx] = var[by][bx];
                    }
                int subs = subsampling < (dec ? 1 : 4) ? subsampling : (dec ? 1 : 4);
                const int damp = damping + cs - (pli != 0);
                for (int g = 0; g < ng; g++) {
                    uint64_t* m = &mse[(size_t)((pli ? 1 : 0) * nfb + fb) * ng + g];
                    const int sv = pli ? str_uv[g] : str_y[g];
                    if (sv < 0) {
                        *m = (uint64_t)1040400 * 64;
                        continue;
                    }
                    const int pri = (sv / 4) << cs;
                    int sec = sv % 4;
                    sec = (sec + (sec == 3)) << cs;
                    uint64_t tot = 0;
                    for (int k = 0; k < cnt; k++) {
                        const int by = lst[k][0], bx = lst[k][1];
                        uint16_t blk[64];
                        const int t = pli ? pri : adjust_strength(pri, var[by][bx]);
                        port_cdef_filter_block(blk, bsz, in + bsz * by * BS + bsz * bx, t, sec, pri ? dir[by][bx] : 0, damp, damp,
                                               dec ? 0 : 3, cs, subs);
                        const uint16_t* sp = src[pli] + (size_t)(fbr * fbs + bsz * by) * sstride[pli] + fbc * fbs + bsz * bx;
                        if (pli == 0) tot += dist8(blk, 8, sp, sstride[0], cs, subs);
                        else
                            for (int i = 0; i < bsz; i += subs)
                                for (int j = 0; j < bsz; j++) {
                                    const int e = (int)sp[i * sstride[pli] + j] - (int)blk[i * bsz + j];
                                    tot += (uint64_t)(e * e);
                                }
                    }
                    const uint64_t v = (tot >> (2 * cs)) * (uint64_t)subs;
                    if (pli == 2) *m += v;
                    else *m = v;
                }
            }
        }
}
