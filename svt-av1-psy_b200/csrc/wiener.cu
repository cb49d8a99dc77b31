// wiener.cu -- K9 separable Wiener filter, K11 Wiener statistics (M = Y^T x, H = Y^T Y) (sm_100a).
//
// Reference behaviour restated:
//   svt_av1_wiener_convolve_add_src_c / svt_av1_highbd_wiener_convolve_add_src_c
//     (Source/Lib/Codec/convolve.c:100-147, 194-237 and the *_hip helpers :57-98, 149-192):
//     8-tap (7 + zero) horizontal pass with add-src, rounding round_0 and clamp to
//     WIENER_CLAMP_LIMIT, then vertical pass with add-src, rounding round_1 and pixel clip.
//   svt_av1_compute_stats_c / _highbd_c (Source/Lib/Codec/restoration_pick.c:659-745):
//     per pixel the wiener_win^2 window of (dgd - avg) is the vector y (column-major), x = src - avg;
//     M[k] += y[k] x, H[k][l] += y[k] y[l]; high bit depth divides by 4 / 16 at the end (truncating).
//
// B200 mapping of the statistics (the one dense contraction on the path): a warp walks DOWN one
// pixel column; lane (a,b), a<=b, owns the 7x7 tile of H that pairs window column a with window
// column b, and keeps the two 7-pixel vertical strips in registers as a sliding window -- 2 loads
// feed 49 multiply-accumulates per pixel.  The 7 diagonal lanes also accumulate M.  Products are
// accumulated in int32 for as many pixels as cannot overflow, then flushed to the int64 totals with
// 64-bit atomics; a finalize kernel mirrors the triangle and applies the bit-depth divider.
#include "common.cuh"
#include "../../include/svt_b200.h"

namespace b200 {

__device__ __forceinline__ int round_pow2_s(int v, int n) { return (v + ((1 << n) >> 1)) >> n; }

// ---------------------------------------------------------------------------------------------
// K9
// ---------------------------------------------------------------------------------------------
template <typename PIX>
__global__ void __launch_bounds__(256)
wiener_convolve_kernel(const PIX* __restrict__ src_base, PIX* __restrict__ dst_base, const SvtB200WienerUnit* __restrict__ units,
                       int n_units, int bd, int round0, int round1, int lbd_rows) {
    __shared__ uint16_t s_src[(64 + 8) * (64 + 8)];
    __shared__ uint16_t s_tmp[(64 + 8) * 64];
    for (int it = blockIdx.x; it < n_units; it += gridDim.x) {
        const SvtB200WienerUnit u = units[it];
        const int w = u.w, h = u.h;
        const PIX* src = src_base + u.src_off;
        PIX*       dst = dst_base + u.dst_off;
        const int sw = w + 8, sh = h + 7;  // rows -3..h+3, cols -3..w+4
        for (int i = threadIdx.x; i < sw * sh; i += blockDim.x) {
            const int r = i / sw, c = i - r * sw;
            // the 8th tap is read by the reference too (multiplied by its coefficient); the column
            // w+4 it touches on the last pixel is part of the caller's extended border
            s_src[r * 72 + c] = (uint16_t)src[(ptrdiff_t)(r - 3) * u.src_stride + (c - 3)];
        }
        __syncthreads();
        const int limit = (1 << (bd + 1 + 7 - round0)) - 1;
        // horizontal: intermediate rows -3..h+3; the low-bit-depth reference computes h+6 rows and
        // zero-fills the last one (convolve.c:113-121)
        const int hrows = lbd_rows ? h + 6 : h + 7;
        for (int i = threadIdx.x; i < sh * w; i += blockDim.x) {
            const int r = i / w, c = i - r * w;
            int v = 0;
            if (r < hrows) {
                int sum = ((int)s_src[r * 72 + c + 3] << 7) + (1 << (bd + 6));
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)s_src[r * 72 + c + k] * (int)u.hfilter[k];
                v = round_pow2_s(sum, round0);
                v = v < 0 ? 0 : (v > limit ? limit : v);
            }
            s_tmp[r * 64 + c] = (uint16_t)v;
        }
        __syncthreads();
        const int pmax = (1 << bd) - 1;
        for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
            const int r = i / w, c = i - r * w;
            int sum = ((int)s_tmp[(r + 3) * 64 + c] << 7) - (1 << (bd + round1 - 1));
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (r + k < sh) sum += (int)s_tmp[(r + k) * 64 + c] * (int)u.vfilter[k];
            int v = round_pow2_s(sum, round1);
            v = v < 0 ? 0 : (v > pmax ? pmax : v);
            dst[(ptrdiff_t)r * u.dst_stride + c] = (PIX)v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K11
// ---------------------------------------------------------------------------------------------
template <typename PIX>
__global__ void stats_avg_kernel(const PIX* __restrict__ dgd_base, const SvtB200StatsItem* __restrict__ items, int n_items,
                                 int* __restrict__ avg_out) {
    __shared__ unsigned long long tot;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const SvtB200StatsItem s = items[it];
        if (threadIdx.x == 0) tot = 0;
        __syncthreads();
        const int w = s.h_end - s.h_start, h = s.v_end - s.v_start;
        unsigned long long acc = 0;
        for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
            const int r = i / w, c = i - r * w;
            acc += dgd_base[s.dgd_off + (ptrdiff_t)(s.v_start + r) * s.dgd_stride + s.h_start + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&tot, acc);
        __syncthreads();
        if (threadIdx.x == 0) avg_out[it] = (int)(tot / (unsigned long long)(w * h));  // find_average (restoration_pick.c)
        __syncthreads();
    }
}

constexpr int kStatsWarps = 8;
constexpr int kStatsMaxParts = 16;  // CTAs cooperating on one restoration unit
constexpr int kStatsTileW = 32, kStatsTileH = 64, kStatsPitch = kStatsTileW + 6 + 2;

// partial layout per (item, part): [0, 49*49) = H (upper-triangle tiles only), [2401, 2450) = M
//
// Data path is always the 7x7 geometry: a 5x5 (3x3) window is the centre of the 7x7 one, so the
// lanes of a narrower window simply own the centred column pairs and only the centred 5 (3) strip
// rows are flushed.  The row loop is unrolled by 7 so that the sliding strips live in a register
// ring with compile-time indices -- the loop body is branch-free straight-line IMADs.
template <typename PIX>
__global__ void __launch_bounds__(kStatsWarps * 32)
stats_accum_kernel(const PIX* __restrict__ dgd_base, const PIX* __restrict__ src_base, const SvtB200StatsItem* __restrict__ items,
                   const int* __restrict__ avg_in, int ctas_per_item, long long* __restrict__ partial, int flush_pixels) {
    __shared__ unsigned long long s_acc[2450];
    __shared__ int16_t s_d[(kStatsTileH + 6) * kStatsPitch];
    __shared__ int16_t s_x[kStatsTileH * kStatsTileW];
    const int it = blockIdx.x / ctas_per_item, part = blockIdx.x % ctas_per_item;
    const SvtB200StatsItem s = items[it];
    const int win = s.wiener_win, off = (7 - win) >> 1, win2 = win * win;  // off: first physical row/column of the window
    const int avg = avg_in[it];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2450; i += blockDim.x) s_acc[i] = 0;
    __syncthreads();
    // lane -> logical column pair (a, b), a <= b < win; idle lanes shadow pair (0,0) and never flush
    int a = 0, b = 0;
    bool live = false;
    {
        int t = lane;
        for (int aa = 0; aa < win && !live; aa++) {
            const int cnt = win - aa;
            if (t < cnt) { a = aa; b = aa + t; live = true; }
            else t -= cnt;
        }
    }
    const bool diag = live && a == b;
    const int pa = a + off, pb = b + off;  // physical columns inside the 7-wide strip
    int hacc[49], macc[7];
#pragma unroll
    for (int i = 0; i < 49; i++) hacc[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) macc[i] = 0;
    const PIX* dgd = dgd_base + s.dgd_off;
    const PIX* src = src_base + s.src_off;
    int pending = 0;
    // int32 partial sums -> the CTA's int64 totals in shared memory (8 warps contend at most)
    auto flush = [&]() {
#pragma unroll
        for (int l1 = 0; l1 < 7; l1++)
#pragma unroll
            for (int l2 = 0; l2 < 7; l2++) {
                const int q1 = l1 - off, q2 = l2 - off;  // logical strip rows
                if (live && q1 >= 0 && q1 < win && q2 >= 0 && q2 < win && hacc[l1 * 7 + l2])
                    atomicAdd(&s_acc[(a * win + q1) * win2 + (b * win + q2)], (unsigned long long)(long long)hacc[l1 * 7 + l2]);
                hacc[l1 * 7 + l2] = 0;
            }
#pragma unroll
        for (int l1 = 0; l1 < 7; l1++) {
            const int q1 = l1 - off;
            if (diag && q1 >= 0 && q1 < win && macc[l1]) atomicAdd(&s_acc[2401 + a * win + q1], (unsigned long long)(long long)macc[l1]);
            macc[l1] = 0;
        }
        pending = 0;
    };
    const int W = s.h_end - s.h_start;
    for (int g = part; g * kStatsTileW < W; g += ctas_per_item) {
        const int c0 = s.h_start + g * kStatsTileW, ncols = min(kStatsTileW, s.h_end - c0);
        for (int r0 = s.v_start; r0 < s.v_end; r0 += kStatsTileH) {
            const int nrows = min(kStatsTileH, s.v_end - r0);
            __syncthreads();
            // rows r0-3 .. r0+nrows+2, columns c0-3 .. c0+ncols+2 as (pixel - avg)
            for (int t = threadIdx.x; t < (nrows + 6) * (ncols + 6); t += blockDim.x) {
                const int rr = t / (ncols + 6), cc = t - rr * (ncols + 6);
                s_d[rr * kStatsPitch + cc] = (int16_t)((int)dgd[(ptrdiff_t)(r0 - 3 + rr) * s.dgd_stride + c0 - 3 + cc] - avg);
            }
            for (int t = threadIdx.x; t < nrows * ncols; t += blockDim.x) {
                const int rr = t / ncols, cc = t - rr * ncols;
                s_x[rr * kStatsTileW + cc] = (int16_t)((int)src[(ptrdiff_t)(r0 + rr) * s.src_stride + c0 + cc] - avg);
            }
            __syncthreads();
            for (int k = 0; k < kStatsTileW / kStatsWarps; k++) {
                const int cw = warp * (kStatsTileW / kStatsWarps) + k;
                if (cw >= ncols) break;  // warp-uniform
                if (pending + nrows > flush_pixels) flush();
                pending += nrows;
                const int16_t* da = s_d + cw + pa;
                const int16_t* db = s_d + cw + pb;
                const int16_t* dx = s_x + cw;
                // register ring: slot (r mod 7) holds tile row r
                int ra[7], rb[7];
#pragma unroll
                for (int l = 0; l < 6; l++) {
                    ra[l] = da[l * kStatsPitch];
                    rb[l] = db[l * kStatsPitch];
                }
                ra[6] = rb[6] = 0;
                int i0 = 0;
                for (; i0 + 7 <= nrows; i0 += 7) {
#pragma unroll
                    for (int t = 0; t < 7; t++) {
                        // pixel row i0+t uses tile rows i0+t .. i0+t+6; the new one goes to slot (t+6) % 7
                        ra[(t + 6) % 7] = da[(i0 + t + 6) * kStatsPitch];
                        rb[(t + 6) % 7] = db[(i0 + t + 6) * kStatsPitch];
                        const int x = dx[(i0 + t) * kStatsTileW];
#pragma unroll
                        for (int l1 = 0; l1 < 7; l1++) {
#pragma unroll
                            for (int l2 = 0; l2 < 7; l2++) hacc[l1 * 7 + l2] += ra[(t + l1) % 7] * rb[(t + l2) % 7];
                            macc[l1] += ra[(t + l1) % 7] * x;
                        }
                    }
                }
                // tail (< 7 rows): same arithmetic with a shifting strip
                if (i0 < nrows) {
                    int ya[7], yb[7];
#pragma unroll
                    for (int l = 0; l < 6; l++) {
                        ya[l + 1] = da[(i0 + l) * kStatsPitch];
                        yb[l + 1] = db[(i0 + l) * kStatsPitch];
                    }
                    ya[0] = yb[0] = 0;
                    for (int i = i0; i < nrows; i++) {
#pragma unroll
                        for (int l = 0; l < 6; l++) {
                            ya[l] = ya[l + 1];
                            yb[l] = yb[l + 1];
                        }
                        ya[6] = da[(i + 6) * kStatsPitch];
                        yb[6] = db[(i + 6) * kStatsPitch];
                        const int x = dx[i * kStatsTileW];
#pragma unroll
                        for (int l1 = 0; l1 < 7; l1++) {
#pragma unroll
                            for (int l2 = 0; l2 < 7; l2++) hacc[l1 * 7 + l2] += ya[l1] * yb[l2];
                            macc[l1] += ya[l1] * x;
                        }
                    }
                }
            }
        }
    }
    flush();
    __syncthreads();
    long long* P = partial + ((size_t)it * ctas_per_item + part) * 2450;
    for (int i = threadIdx.x; i < 2450; i += blockDim.x) P[i] = (long long)s_acc[i];
}

__global__ void stats_finalize_kernel(const long long* __restrict__ partial, int parts, const SvtB200StatsItem* __restrict__ items,
                                      int n_items, int divider, long long* __restrict__ M_out, long long* __restrict__ H_out) {
    __shared__ long long A[2450];
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int win = items[it].wiener_win, win2 = win * win;
        __syncthreads();
        for (int i = threadIdx.x; i < 2450; i += blockDim.x) {
            long long v = 0;
            for (int p = 0; p < parts; p++) v += partial[((size_t)it * parts + p) * 2450 + i];
            A[i] = v;
        }
        __syncthreads();
        long long* M = M_out + (size_t)it * 49;
        long long* H = H_out + (size_t)it * 2401;
        for (int k = threadIdx.x; k < win2; k += blockDim.x) M[k] = A[2401 + k] / divider;
        for (int p = threadIdx.x; p < win2 * win2; p += blockDim.x) {
            const int k = p / win2, l = p - k * win2;
            // tiles were accumulated for window-column pairs a<=b only: element (k,l) lives in the
            // tile of (k/win, l/win) when k/win <= l/win, else in its mirror
            const int ka = k / win, la = l / win;
            const long long v = (ka <= la) ? A[k * win2 + l] : A[l * win2 + k];
            H[p] = v / divider;
        }
    }
}

static long long* g_stats_acc = nullptr;
static int*       g_stats_avg = nullptr;
static size_t     g_stats_cap = 0;
static std::mutex g_stats_mu;

template <typename PIX>
static void launch_stats(const PIX* d_dgd, const PIX* d_src, const SvtB200StatsItem* d_items, int n, int bd, long long* d_M,
                         long long* d_H, long long* d_acc, int* d_avg, cudaStream_t st) {
    const int maxv = (1 << bd) - 1;
    long long fp = 2147483647ll / ((long long)maxv * maxv);
    if (fp > 30000) fp = 30000;
    if (fp < 1) fp = 1;
    const int divider = bd == 12 ? 16 : (bd == 10 ? 4 : 1);
    int cpi = (ctx().sm_count * 4) / (n > 0 ? n : 1);
    if (cpi < 1) cpi = 1;
    if (cpi > kStatsMaxParts) cpi = kStatsMaxParts;
    stats_avg_kernel<PIX><<<grid_for(n, 4), 256, 0, st>>>(d_dgd, d_items, n, d_avg);
    B200_LAUNCH_CHECK();
    stats_accum_kernel<PIX><<<n * cpi, kStatsWarps * 32, 0, st>>>(d_dgd, d_src, d_items, d_avg, cpi, d_acc, (int)fp);
    B200_LAUNCH_CHECK();
    stats_finalize_kernel<<<grid_for(n, 4), 256, 0, st>>>(d_acc, cpi, d_items, n, divider, d_M, d_H);
    B200_LAUNCH_CHECK();
}

template <typename PIX>
static void stats_t1(int wiener_win, const PIX* dgd, const PIX* src, int h_start, int h_end, int v_start, int v_end, int dgd_stride,
                     int src_stride, int64_t* M, int64_t* H, int bd) {
    require_ready();
    const int half = wiener_win >> 1, win2 = wiener_win * wiener_win;
    const int w = h_end - h_start, h = v_end - v_start;
    const int dw = w + 2 * half, dh = h + 2 * half;
    LaneGuard l;
    size_t o_d = l->alloc((size_t)dw * dh * sizeof(PIX)), o_s = l->alloc((size_t)w * h * sizeof(PIX)), o_it = l->alloc(sizeof(SvtB200StatsItem));
    size_t in_end = l->used;
    size_t o_M = l->alloc(49 * 8), o_H = l->alloc(2401 * 8), o_acc = l->alloc((size_t)kStatsMaxParts * 2450 * 8), o_avg = l->alloc(16);
    for (int r = 0; r < dh; r++)
        memcpy(l->h<PIX>(o_d) + (size_t)r * dw, dgd + (ptrdiff_t)(v_start - half + r) * dgd_stride + h_start - half, dw * sizeof(PIX));
    for (int r = 0; r < h; r++) memcpy(l->h<PIX>(o_s) + (size_t)r * w, src + (ptrdiff_t)(v_start + r) * src_stride + h_start, w * sizeof(PIX));
    SvtB200StatsItem* it = l->h<SvtB200StatsItem>(o_it);
    memset(it, 0, sizeof(*it));
    it->dgd_off = (uint64_t)half * dw + half;  // (v_start, h_start) of the packed copy
    it->src_off = 0;
    it->dgd_stride = dw;
    it->src_stride = w;
    it->h_start = 0;
    it->h_end = w;
    it->v_start = 0;
    it->v_end = h;
    it->wiener_win = wiener_win;
    l->h2d(0, in_end);
    launch_stats<PIX>(l->d<PIX>(o_d), l->d<PIX>(o_s), l->d<SvtB200StatsItem>(o_it), 1, bd, l->d<long long>(o_M), l->d<long long>(o_H),
                      l->d<long long>(o_acc), l->d<int>(o_avg), l->stream);
    l->d2h(o_M, (o_H + 2401 * 8) - o_M);
    l->sync();
    memcpy(M, l->h<int64_t>(o_M), (size_t)win2 * 8);
    memcpy(H, l->h<int64_t>(o_H), (size_t)win2 * win2 * 8);
}

template <typename PIX>
static void wiener_t1(const PIX* src, ptrdiff_t src_stride, PIX* dst, ptrdiff_t dst_stride, const int16_t* fx, const int16_t* fy, int w,
                      int h, int round0, int round1, int bd, int lbd_rows) {
    require_ready();
    LaneGuard l;
    const int sw = w + 8, sh = h + 7;
    size_t o_s = l->alloc((size_t)sw * sh * sizeof(PIX)), o_u = l->alloc(sizeof(SvtB200WienerUnit));
    size_t in_end = l->used;
    size_t o_d = l->alloc((size_t)w * h * sizeof(PIX));
    for (int r = 0; r < sh; r++) memcpy(l->h<PIX>(o_s) + (size_t)r * sw, src + (ptrdiff_t)(r - 3) * src_stride - 3, sw * sizeof(PIX));
    SvtB200WienerUnit* u = l->h<SvtB200WienerUnit>(o_u);
    memset(u, 0, sizeof(*u));
    u->src_off = (uint64_t)3 * sw + 3;
    u->dst_off = 0;
    u->src_stride = sw;
    u->dst_stride = w;
    u->w = (uint16_t)w;
    u->h = (uint16_t)h;
    memcpy(u->hfilter, fx, 16);
    memcpy(u->vfilter, fy, 16);
    l->h2d(0, in_end);
    wiener_convolve_kernel<PIX><<<1, 256, 0, l->stream>>>(l->d<PIX>(o_s), l->d<PIX>(o_d), l->d<SvtB200WienerUnit>(o_u), 1, bd, round0, round1, lbd_rows);
    B200_LAUNCH_CHECK();
    l->d2h(o_d, (size_t)w * h * sizeof(PIX));
    l->sync();
    for (int r = 0; r < h; r++) memcpy(dst + (ptrdiff_t)r * dst_stride, l->h<PIX>(o_d) + (size_t)r * w, w * sizeof(PIX));
}

}  // namespace b200

using namespace b200;

extern "C" void svt_b200_av1_wiener_convolve_add_src(const uint8_t* src, ptrdiff_t src_stride, uint8_t* dst, ptrdiff_t dst_stride,
                                                     const int16_t* filter_x, const int16_t* filter_y, int32_t w, int32_t h,
                                                     const SvtB200ConvolveParams* conv_params) {
    wiener_t1<uint8_t>(src, src_stride, dst, dst_stride, filter_x, filter_y, w, h, conv_params->round_0, conv_params->round_1, 8, 1);
}
extern "C" void svt_b200_av1_highbd_wiener_convolve_add_src(const uint16_t* src, ptrdiff_t src_stride, uint16_t* dst,
                                                            ptrdiff_t dst_stride, const int16_t* filter_x, const int16_t* filter_y,
                                                            int32_t w, int32_t h, const SvtB200ConvolveParams* conv_params, int32_t bd) {
    wiener_t1<uint16_t>(src, src_stride, dst, dst_stride, filter_x, filter_y, w, h, conv_params->round_0, conv_params->round_1, bd, 0);
}
extern "C" void svt_b200_av1_compute_stats(int32_t wiener_win, const uint8_t* dgd, const uint8_t* src, int32_t h_start, int32_t h_end,
                                           int32_t v_start, int32_t v_end, int32_t dgd_stride, int32_t src_stride, int64_t* M, int64_t* H) {
    stats_t1<uint8_t>(wiener_win, dgd, src, h_start, h_end, v_start, v_end, dgd_stride, src_stride, M, H, 8);
}
extern "C" void svt_b200_av1_compute_stats_highbd(int32_t wiener_win, const uint16_t* dgd, const uint16_t* src, int32_t h_start,
                                                  int32_t h_end, int32_t v_start, int32_t v_end, int32_t dgd_stride, int32_t src_stride,
                                                  int64_t* M, int64_t* H, int32_t bit_depth) {
    stats_t1<uint16_t>(wiener_win, dgd, src, h_start, h_end, v_start, v_end, dgd_stride, src_stride, M, H, bit_depth);
}

extern "C" int svt_b200_wiener_units_dev(const void* d_src, void* d_dst, const SvtB200WienerUnit* d_units, int n_units, int bit_depth,
                                         void* stream) {
    require_ready();
    if (n_units <= 0) return n_units == 0 ? SVT_B200_OK : SVT_B200_ERR_BAD_ARG;
    // get_conv_params_wiener (convolve.h): round_0 = 3 (+2 at 12 bit), round_1 = 2*FILTER_BITS - round_0
    const int round0 = bit_depth == 12 ? 5 : 3, round1 = 14 - round0;
    cudaStream_t st = (cudaStream_t)stream;
    if (bit_depth > 8)
        wiener_convolve_kernel<uint16_t><<<grid_for(n_units, 4), 256, 0, st>>>((const uint16_t*)d_src, (uint16_t*)d_dst, d_units, n_units, bit_depth, round0, round1, 0);
    else
        wiener_convolve_kernel<uint8_t><<<grid_for(n_units, 4), 256, 0, st>>>((const uint8_t*)d_src, (uint8_t*)d_dst, d_units, n_units, 8, round0, round1, 1);
    B200_LAUNCH_CHECK();
    return SVT_B200_OK;
}

extern "C" int svt_b200_compute_stats_batch_dev(const void* d_dgd, const void* d_src, const SvtB200StatsItem* d_items, int n_items,
                                                int bit_depth, int64_t* d_M, int64_t* d_H, void* stream) {
    require_ready();
    if (n_items <= 0) return n_items == 0 ? SVT_B200_OK : SVT_B200_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(g_stats_mu);
    if ((size_t)n_items > g_stats_cap) {
        if (g_stats_acc) { cudaFree(g_stats_acc); cudaFree(g_stats_avg); }
        g_stats_cap = (size_t)n_items * 2;
        B200_CUDA_CHECK(cudaMalloc(&g_stats_acc, g_stats_cap * kStatsMaxParts * 2450 * 8));
        B200_CUDA_CHECK(cudaMalloc(&g_stats_avg, g_stats_cap * 4));
    }
    if (bit_depth > 8)
        launch_stats<uint16_t>((const uint16_t*)d_dgd, (const uint16_t*)d_src, d_items, n_items, bit_depth, (long long*)d_M, (long long*)d_H,
                               g_stats_acc, g_stats_avg, (cudaStream_t)stream);
    else
        launch_stats<uint8_t>((const uint8_t*)d_dgd, (const uint8_t*)d_src, d_items, n_items, 8, (long long*)d_M, (long long*)d_H, g_stats_acc,
                              g_stats_avg, (cudaStream_t)stream);
    return SVT_B200_OK;
}
