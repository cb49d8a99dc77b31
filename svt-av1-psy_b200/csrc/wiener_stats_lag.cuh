// wiener_stats_lag.cuh -- K11 Wiener statistics by LAG sums (sm_100a), included by wiener.cu.
//
// Reference: svt_av1_compute_stats_c / _highbd_c (Source/Lib/Codec/restoration_pick.c:659-745).  With y_p(i,j) =
// dgd(i + lr_p, j + kc_p) - avg the window sample p of pixel (i,j) (p = (kc + half) * win + (lr + half)) and x = src - avg,
//     M[p] = sum_unit x y_p          H[p][q] = sum_unit y_p y_q .
// Brute force costs win^2 (win^2 + 1) / 2 + win^2 = 1274 multiply-accumulates per pixel (7x7).  But y_p y_q only depends on
// the LAG (dy, dx) = (lr_q - lr_p, kc_q - kc_p): H[p][q] is the sum of the lag's product image P(u,v) = Y(u,v) Y(u+dy, v+dx)
// over the unit rectangle shifted by (lr_p, kc_p).  All win^2 shifted rectangles of a lag share the CORE
// [vs+half, ve-half) x [hs+half, he-half) and differ in a frame of 2*half rows / columns around it.  So, exactly, in integers:
//     H[p][q] = CC(lag) + sum of the edge-row sums RS(lag, u) that the shifted rectangle contains
//                       + sum of the edge-column sums CS(lag, v) ...   + the <= (2 half)^2 corner products,
// with RAW pixels (no mean removed: products are non-negative, 8-bit pictures take four multiply-accumulates per DP4A) and
// the mean folded back in at the end:  sum (y-a)(z-a) = sum yz - a (sum y + sum z) + N a^2  (a = the integer average).
// That is (2 win - 1) win - (win - 1) = 85 lags + 49 cross lags for M per pixel instead of 1274 -- an order of magnitude
// fewer operations than the tensor-core formulation needs, identical results, same code for 8 / 10 / 12 bit.
//
//   stats_lag_bulk_kernel   CC of every lag, sum x Y(shift) for M, plain pixel sums: one CTA per (8-row band, unit); warp g owns
//                           the lags with dy = g and the M shifts with lr = g - half; partner words come from a 5 / 7 word
//                           window per row with compile-time funnel shifts.
//   stats_lag_edges_kernel  RS / CS: the 4 half edge rows and columns of every lag.
//   stats_lag_finalize      assembles H (both triangles) and M, applies the mean correction and the bit-depth divider.
#pragma once

namespace b200 {

constexpr int kLagSlots = 92;                 // dy * 13 + (dx + 6) for the 7x7 geometry, slot 91 = "ones" (plain pixel sums)
constexpr int kLagOnes = 91;
constexpr int kLagAccStride = 144;            // per item: [0,92) CC, [92,141) sum x Y(shift p), [141] sum x
constexpr int kLagEdge = 12;                  // edge rows (columns) per lag: 2 groups of 2 * half <= 6
constexpr int kLagItemWords = kLagAccStride + 2 * kLagSlots * kLagEdge;  // + RS + CS  (uint64 each)
constexpr int kLagBandRows = 8;

template <typename PIX> struct LagPix;
template <> struct LagPix<uint8_t> {
    static constexpr int G = 4, OFFP = 8;     // pixels per 32-bit word; pixels between the tile row origin and the first core column
    static constexpr uint32_t ONES = 0x01010101u;
    static __device__ __forceinline__ uint32_t mac(uint32_t a, uint32_t b, uint32_t acc) { return __dp4a(a, b, acc); }
};
template <> struct LagPix<uint16_t> {
    static constexpr int G = 2, OFFP = 6;
    static constexpr uint32_t ONES = 0x00010001u;
    static __device__ __forceinline__ uint32_t mac(uint32_t a, uint32_t b, uint32_t acc) {
        return acc + (a & 0xffffu) * (b & 0xffffu) + (a >> 16) * (b >> 16);
    }
};

// edge rows of a unit: slot i < 2 half -> vs - half + i; slot 2 half + i -> max(ve - half, vs + half) + i (invalid when >= ve + half)
__host__ __device__ __forceinline__ int lag_edge_line(int lo, int hi, int half, int i) {
    if (i < 2 * half) return lo - half + i;
    const int b0 = (hi - half) > (lo + half) ? (hi - half) : (lo + half);
    const int l = b0 + (i - 2 * half);
    return l < hi + half ? l : 0x7fffffff;
}

constexpr int kLagSegW = 512;  // pixels per column segment of a work unit (any unit width is handled: segments are additive)

// grid = (CTAs per unit, units); a CTA walks the (row band, column segment) work units of its restoration unit
template <typename PIX, int WIN>
__global__ void __launch_bounds__(256)
stats_lag_bulk_kernel(const PIX* __restrict__ dgd_base, const PIX* __restrict__ src_base, const SvtB200StatsItem* __restrict__ items,
                      unsigned long long* __restrict__ acc_base) {
    using P = LagPix<PIX>;
    constexpr int HALF = WIN >> 1, S = (int)sizeof(PIX), G = P::G, OFFB = P::OFFP * S;
    constexpr int NW = (OFFB + (WIN - 1) * S + 3) / 4 + 1;  // words of the partner window of one core word (5 for 8-bit, 7 for 16-bit at 7x7)
    constexpr int PITCH = (OFFB + (kLagSegW + 2 * (WIN - 1)) * S + 3) / 4 + NW + 1;  // words per Y tile row
    constexpr int XPITCH = (kLagSegW * S + 3) / 4 + 1;
    constexpr int TROWS = kLagBandRows + 3 * HALF;
    const SvtB200StatsItem s = items[blockIdx.y];
    if (s.wiener_win != WIN) return;
    const int hs = s.h_start, he = s.h_end, vs = s.v_start, ve = s.v_end, w = he - hs, h = ve - vs;
    const int cu0 = vs + HALF, cu1 = ve - HALF, cv0 = hs + HALF, cv1 = he - HALF;
    extern __shared__ __align__(16) uint32_t lag_sm[];
    uint32_t* Yt = lag_sm;                  // rows [b0 - HALF, b1 + 2 HALF), row byte 0 = pixel column cv0 + c0 - OFFP; outside the halo'd unit: 0
    uint32_t* Xt = lag_sm + TROWS * PITCH;  // rows [b0, b1) of the source, word 0 = column hs + c0
    const PIX* dgd = dgd_base + s.dgd_off;
    const PIX* src = src_base + s.src_off;
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long totH[2 * WIN - 1], totM[WIN], totY = 0, totX = 0;
#pragma unroll
    for (int k = 0; k < 2 * WIN - 1; k++) totH[k] = 0;
#pragma unroll
    for (int k = 0; k < WIN; k++) totM[k] = 0;
    const int nbands = (h + kLagBandRows - 1) / kLagBandRows, nsegs = (w + kLagSegW - 1) / kLagSegW;
    for (int wu = blockIdx.x; wu < nbands * nsegs; wu += gridDim.x) {
        const int band = wu / nsegs, seg = wu - band * nsegs;
        const int b0 = vs + band * kLagBandRows, b1 = min(b0 + kLagBandRows, ve), c0 = seg * kLagSegW;
        const int cw = max(min(cv1 - (cv0 + c0), kLagSegW), 0);  // core columns of this segment
        const int xw = min(w - c0, kLagSegW);                      // unit columns of this segment
        __syncthreads();
        {
            PIX* yb = reinterpret_cast<PIX*>(Yt);
            constexpr int PPR = PITCH * 4 / S;  // pixels per tile row
            const int trows = (b1 - b0) + 3 * HALF;
            for (int i = threadIdx.x; i < trows * PPR; i += blockDim.x) {
                const int tr = i / PPR, tc = i - tr * PPR;
                const int u = b0 - HALF + tr, v = cv0 + c0 - P::OFFP + tc;
                const bool in = u >= vs - HALF && u < ve + HALF && v >= hs - HALF && v < he + HALF;
                yb[i] = in ? dgd[(ptrdiff_t)u * s.dgd_stride + v] : (PIX)0;
            }
            PIX* xb = reinterpret_cast<PIX*>(Xt);
            constexpr int XPR = XPITCH * 4 / S;
            for (int i = threadIdx.x; i < (b1 - b0) * XPR; i += blockDim.x) {
                const int tr = i / XPR, tc = i - tr * XPR;
                xb[i] = tc < xw ? src[(ptrdiff_t)(b0 + tr) * s.src_stride + hs + c0 + tc] : (PIX)0;
            }
        }
        __syncthreads();
        uint32_t accH[2 * WIN - 1], accM[WIN], accY = 0, accX = 0;
#pragma unroll
        for (int k = 0; k < 2 * WIN - 1; k++) accH[k] = 0;
#pragma unroll
        for (int k = 0; k < WIN; k++) accM[k] = 0;
        const int ncw = (cw + G - 1) / G, nxw = (xw + G - 1) / G;  // words of a core row / of a unit row of this segment
        if (g < WIN) {
            // ---- H lags with dy = g: A = core word of row u, partner words from row u + g ---------------------------------------
            const int dy = g;
            for (int u = max(b0, cu0); u < min(b1, cu1); u++) {
                const uint32_t* arow = Yt + (u - (b0 - HALF)) * PITCH + OFFB / 4;
                const uint32_t* brow = Yt + (u + dy - (b0 - HALF)) * PITCH;
                for (int wc = lane; wc < ncw; wc += 32) {
                    uint32_t a = arow[wc];
                    const int rem = cw - wc * G;  // pixels of this word inside the core
                    if (rem < G) a &= (1u << (rem * 8 * S)) - 1u;
                    uint32_t win_w[NW];
#pragma unroll
                    for (int k = 0; k < NW; k++) win_w[k] = brow[wc + k];
#pragma unroll
                    for (int k = 0; k < 2 * WIN - 1; k++) {
                        const int dx = k - (WIN - 1);
                        const int rel = OFFB + dx * S;  // byte offset of the partner inside the window (compile-time after unrolling)
                        const uint32_t b = (rel & 3) ? __funnelshift_r(win_w[rel >> 2], win_w[(rel >> 2) + 1], (rel & 3) * 8) : win_w[rel >> 2];
                        if (dy > 0 || dx >= 0) accH[k] = P::mac(a, b, accH[k]);
                    }
                }
            }
            // ---- M shifts with lr = g - HALF: A = source word of unit row u, partners = Y(u + lr, column + kc) -------------------
            const int lr = g - HALF;
            for (int u = b0; u < b1; u++) {
                const uint32_t* arow = Xt + (u - b0) * XPITCH;
                const uint32_t* brow = Yt + (u + lr - (b0 - HALF)) * PITCH;
                for (int wc = lane; wc < nxw; wc += 32) {
                    const uint32_t a = arow[wc];  // columns >= xw were staged as 0
                    uint32_t win_w[NW];
#pragma unroll
                    for (int k = 0; k < NW; k++) win_w[k] = brow[wc + k];
#pragma unroll
                    for (int k = 0; k < WIN; k++) {
                        const int kc = k - HALF;
                        const int rel = (P::OFFP - HALF + kc) * S;
                        const uint32_t b = (rel & 3) ? __funnelshift_r(win_w[rel >> 2], win_w[(rel >> 2) + 1], (rel & 3) * 8) : win_w[rel >> 2];
                        accM[k] = P::mac(a, b, accM[k]);
                    }
                }
            }
        } else if (g == WIN) {
            // ---- plain sums: Y over the core rows of the band ("ones" lag), x over the unit rows -----------------------------------
            for (int u = max(b0, cu0); u < min(b1, cu1); u++) {
                const uint32_t* arow = Yt + (u - (b0 - HALF)) * PITCH + OFFB / 4;
                for (int wc = lane; wc < ncw; wc += 32) {
                    uint32_t a = arow[wc];
                    const int rem = cw - wc * G;
                    if (rem < G) a &= (1u << (rem * 8 * S)) - 1u;
                    accY = P::mac(a, P::ONES, accY);
                }
            }
            for (int u = b0; u < b1; u++) {
                const uint32_t* arow = Xt + (u - b0) * XPITCH;
                for (int wc = lane; wc < nxw; wc += 32) accX = P::mac(arow[wc], P::ONES, accX);
            }
        }
        // a work unit's 32-bit partial sums cannot overflow (8 rows x 512 columns / 32 lanes of <= 2 x 4095^2); across work units: 64 bit
#pragma unroll
        for (int k = 0; k < 2 * WIN - 1; k++) totH[k] += accH[k];
#pragma unroll
        for (int k = 0; k < WIN; k++) totM[k] += accM[k];
        totY += accY;
        totX += accX;
    }
    unsigned long long* acc = acc_base + (size_t)blockIdx.y * kLagItemWords;
    auto flush = [&](unsigned long long t, int slot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && t) atomicAdd(&acc[slot], t);
    };
    if (g < WIN) {
#pragma unroll
        for (int k = 0; k < 2 * WIN - 1; k++)
            if (g > 0 || k >= WIN - 1) flush(totH[k], g * 13 + (k - (WIN - 1)) + 6);
#pragma unroll
        for (int k = 0; k < WIN; k++) flush(totM[k], kLagSlots + k * WIN + g);  // p = (kc + half) * win + (lr + half)
    } else if (g == WIN) {
        flush(totY, kLagOnes);
        flush(totX, kLagSlots + 49);
    }
}

template <typename PIX, int WIN>
constexpr size_t lag_bulk_smem() {
    using P = LagPix<PIX>;
    constexpr int S = (int)sizeof(PIX), OFFB = P::OFFP * S, NW = (OFFB + (WIN - 1) * S + 3) / 4 + 1;
    constexpr int PITCH = (OFFB + (kLagSegW + 2 * (WIN - 1)) * S + 3) / 4 + NW + 1, XPITCH = (kLagSegW * S + 3) / 4 + 1;
    return (size_t)((kLagBandRows + 3 * (WIN >> 1)) * PITCH + kLagBandRows * XPITCH) * 4;
}

// RS(lag, edge row u) = sum over the core columns of P(u, v); CS(lag, edge column v) = sum over the core rows.  grid = (92, items)
template <typename PIX>
__global__ void __launch_bounds__(256)
stats_lag_edges_kernel(const PIX* __restrict__ dgd_base, const SvtB200StatsItem* __restrict__ items, unsigned long long* __restrict__ acc_base) {
    const SvtB200StatsItem s = items[blockIdx.y];
    const int win = s.wiener_win, half = win >> 1, slot = blockIdx.x;
    const bool ones = slot == kLagOnes;
    const int dy = ones ? 0 : slot / 13, dx = ones ? 0 : slot % 13 - 6;
    if (!ones && (dy >= win || dx <= -win || dx >= win || (dy == 0 && dx < 0))) return;
    const int hs = s.h_start, he = s.h_end, vs = s.v_start, ve = s.v_end;
    const int cu0 = vs + half, cu1 = ve - half, cv0 = hs + half, cv1 = he - half;
    const PIX* dgd = dgd_base + s.dgd_off;
    unsigned long long* out = acc_base + (size_t)blockIdx.y * kLagItemWords + kLagAccStride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int line = warp; line < 2 * kLagEdge; line += 8) {
        const bool is_row = line < kLagEdge;
        const int i = is_row ? line : line - kLagEdge;
        unsigned long long t = 0;
        if (i < 4 * half) {
            if (is_row) {
                const int u = lag_edge_line(vs, ve, half, i);
                if (u != 0x7fffffff && u + dy < ve + half)
                    for (int v = cv0 + lane; v < cv1; v += 32) {
                        const unsigned long long a = dgd[(ptrdiff_t)u * s.dgd_stride + v];
                        t += ones ? a : a * (unsigned long long)dgd[(ptrdiff_t)(u + dy) * s.dgd_stride + v + dx];
                    }
            } else {
                const int v = lag_edge_line(hs, he, half, i);
                if (v != 0x7fffffff && v + dx >= hs - half && v + dx < he + half)
                    for (int u = cu0 + lane; u < cu1; u += 32) {
                        const unsigned long long a = dgd[(ptrdiff_t)u * s.dgd_stride + v];
                        t += ones ? a : a * (unsigned long long)dgd[(ptrdiff_t)(u + dy) * s.dgd_stride + v + dx];
                    }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) out[(size_t)(is_row ? 0 : kLagSlots * kLagEdge) + slot * kLagEdge + i] = t;
    }
}

// sum of a lag's product image over the unit rectangle shifted by (ar, ac): core + contained edge rows / columns + corners
template <typename PIX>
__device__ long long lag_rect_sum(const PIX* dgd, const SvtB200StatsItem& s, const unsigned long long* acc, int slot, bool ones, int dy, int dx,
                                  int ar, int ac) {
    const int half = s.wiener_win >> 1, hs = s.h_start, he = s.h_end, vs = s.v_start, ve = s.v_end;
    const unsigned long long* RS = acc + kLagAccStride + slot * kLagEdge;
    const unsigned long long* CS = acc + kLagAccStride + kLagSlots * kLagEdge + slot * kLagEdge;
    long long tot = (long long)acc[slot];
    for (int i = 0; i < 4 * half; i++) {
        const int u = lag_edge_line(vs, ve, half, i);
        if (u != 0x7fffffff && u >= vs + ar && u < ve + ar) tot += (long long)RS[i];
        const int v = lag_edge_line(hs, he, half, i);
        if (v != 0x7fffffff && v >= hs + ac && v < he + ac) tot += (long long)CS[i];
    }
    for (int i = 0; i < 4 * half; i++) {
        const int u = lag_edge_line(vs, ve, half, i);
        if (u == 0x7fffffff || u < vs + ar || u >= ve + ar) continue;
        for (int j = 0; j < 4 * half; j++) {
            const int v = lag_edge_line(hs, he, half, j);
            if (v == 0x7fffffff || v < hs + ac || v >= he + ac) continue;
            const long long a = dgd[(ptrdiff_t)u * s.dgd_stride + v];
            tot += ones ? a : a * (long long)dgd[(ptrdiff_t)(u + dy) * s.dgd_stride + v + dx];
        }
    }
    return tot;
}

// grid = (ceil((win2 (win2 + 1) / 2 + win2) / 128), items)
template <typename PIX>
__global__ void __launch_bounds__(128)
stats_lag_finalize_kernel(const PIX* __restrict__ dgd_base, const SvtB200StatsItem* __restrict__ items, const unsigned long long* __restrict__ acc_base,
                          const unsigned long long* __restrict__ tot, int divider, long long* __restrict__ M_out, long long* __restrict__ H_out) {
    const int it = blockIdx.y;
    const SvtB200StatsItem s = items[it];
    const int win = s.wiener_win, half = win >> 1, win2 = win * win;
    const int npairs = win2 * (win2 + 1) / 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= npairs + win2) return;
    const unsigned long long* acc = acc_base + (size_t)it * kLagItemWords;
    const PIX* dgd = dgd_base + s.dgd_off;
    const long long N = (long long)(s.h_end - s.h_start) * (s.v_end - s.v_start);
    const long long avg = (long long)(tot[it] / (unsigned long long)N);  // find_average
    if (e >= npairs) {  // M[p] = sum x y_p - a sum x - a sum Y_p + N a^2
        const int p = e - npairs, kc = p / win - half, lr = p % win - half;
        const long long sy = lag_rect_sum<PIX>(dgd, s, acc, kLagOnes, true, 0, 0, lr, kc);
        const long long m = (long long)acc[kLagSlots + p] - avg * (long long)acc[kLagSlots + 49] - avg * sy + N * avg * avg;
        M_out[(size_t)it * 49 + p] = m / divider;
        return;
    }
    // e -> (p, q), p <= q, rows of the upper triangle in order
    int p = 0, rem = e;
    while (rem >= win2 - p) { rem -= win2 - p; p++; }
    const int q = p + rem;
    const int kp = p / win - half, lp = p % win - half, kq = q / win - half, lq = q % win - half;
    int dy = lq - lp, dx = kq - kp, ar = lp, ac = kp;
    if (dy < 0 || (dy == 0 && dx < 0)) { dy = -dy; dx = -dx; ar = lq; ac = kq; }
    const long long syy = lag_rect_sum<PIX>(dgd, s, acc, dy * 13 + dx + 6, false, dy, dx, ar, ac);
    const long long sp = lag_rect_sum<PIX>(dgd, s, acc, kLagOnes, true, 0, 0, lp, kp);
    const long long sq = p == q ? sp : lag_rect_sum<PIX>(dgd, s, acc, kLagOnes, true, 0, 0, lq, kq);
    const long long hv = (syy - avg * (sp + sq) + N * avg * avg) / divider;
    H_out[(size_t)it * 2401 + p * win2 + q] = hv;
    H_out[(size_t)it * 2401 + q * win2 + p] = hv;
}

}  // namespace b200
