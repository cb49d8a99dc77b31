// me_pyramid.cu -- K2 SAD pyramid (8x8 -> 16x16 -> 32x32 -> 64x64) and the T2 full-pel search that
// is built from it (sm_100a).
//
// Reference behaviour restated:
//   svt_ext_all_sad_calculation_8x8_16x16_c   Source/Lib/Codec/motion_estimation.c:335-363 (+:210-333)
//   svt_ext_eight_sad_calculation_32x32_64x64_c                                   :369-427
//   svt_ext_sad_calculation_8x8_16x16_c / _32x32_64x64_c (1-point variants)        :98-205
//   open_loop_me_fullpel_search_sblock                                            :781-817
// Block numbering: 16x16 blocks are indexed in 32x32-quadrant order (offsets[] table at :341), the
// four 8x8 of 16x16 p are 4p..4p+3 in raster order.  Every best-SAD update is a strict '<' in search
// order (y outer, x inner), i.e. "first minimum in raster order" per PU -- reproduced here as the
// minimum of the 64-bit key (sad << 32 | raster index).  Result layout = me_context.h:54-75
// (64x64 at 0, 32x32 at 1..4, 16x16 at 5..20, 8x8 at 21..84).
#include "common.cuh"
#include "../../include/svt_b200.h"

namespace b200 {

__device__ __forceinline__ int z16_of(int y16, int x16) { return 4 * (2 * (y16 >> 1) + (x16 >> 1)) + 2 * (y16 & 1) + (x16 & 1); }
__device__ __forceinline__ uint32_t pack_mv(int x, int y) { return ((uint32_t)(y & 0xffff) << 16) | (uint32_t)(x & 0xffff); }

// ---------------------------------------------------------------------------------------------
// T1 kernels (caller state in/out, one launch per reference call)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sad8x8_global(const uint8_t* s, uint32_t ss, const uint8_t* r, uint32_t rs, bool sub) {
    uint32_t acc = 0;
    const int step = sub ? 2 : 1;
    for (int y = 0; y < 8; y += step)
        for (int x = 0; x < 8; x++) {
            int d = (int)s[y * ss + x] - (int)r[y * rs + x];
            acc += (uint32_t)(d < 0 ? -d : d);
        }
    return sub ? acc << 1 : acc;
}

// 64 threads: thread = (16x16 raster index, 8x8 k)
__global__ void ext_all_sad_kernel(const uint8_t* src, uint32_t ss, const uint8_t* ref, uint32_t rs, uint32_t mv,
                                   uint32_t* best8, uint32_t* best16, uint32_t* mv8, uint32_t* mv16, uint32_t* eight16,
                                   int sub) {
    __shared__ uint32_t s8[64][8];
    const int t = threadIdx.x, b16 = t >> 2, k = t & 3;
    const int y16 = b16 >> 2, x16 = b16 & 3, p16 = z16_of(y16, x16);
    const uint8_t* s = src + (16 * y16 + 8 * (k >> 1)) * ss + 16 * x16 + 8 * (k & 1);
    const uint8_t* r = ref + (16 * y16 + 8 * (k >> 1)) * rs + 16 * x16 + 8 * (k & 1);
    const int xmv = (int16_t)(mv & 0xffff), ymv = (int16_t)(mv >> 16);
    uint32_t b = best8[4 * p16 + k], bm = mv8[4 * p16 + k];
    for (int i = 0; i < 8; i++) {
        const uint32_t v = sad8x8_global(s, ss, r + i, rs, sub);
        s8[4 * p16 + k][i] = v;
        if (v < b) {
            b  = v;
            bm = pack_mv((int16_t)(xmv + i), ymv);
        }
    }
    best8[4 * p16 + k] = b;
    mv8[4 * p16 + k]   = bm;
    __syncthreads();
    if (t < 16) {
        uint32_t bb = best16[t], bbm = mv16[t];
        for (int i = 0; i < 8; i++) {
            const uint32_t v = s8[4 * t][i] + s8[4 * t + 1][i] + s8[4 * t + 2][i] + s8[4 * t + 3][i];
            eight16[t * 8 + i] = v;
            if (v < bb) {
                bb  = v;
                bbm = pack_mv((int16_t)(xmv + i), ymv);
            }
        }
        best16[t] = bb;
        mv16[t]   = bbm;
    }
}

// n_idx = 8 (eight-point form, sad16 is [16][8]) or 1 (one-point form, sad16 is [16])
__global__ void ext_sad_32_64_kernel(const uint32_t* sad16, uint32_t* best32, uint32_t* best64, uint32_t* mv32,
                                     uint32_t* mv64, uint32_t mv, uint32_t* sad32, int n_idx) {
    if (threadIdx.x != 0) return;
    const int xmv = (int16_t)(mv & 0xffff), ymv = (int16_t)(mv >> 16);
    for (int i = 0; i < n_idx; i++) {
        uint32_t tot = 0;
        const uint32_t m = n_idx == 8 ? pack_mv((int16_t)(xmv + i), ymv) : mv;
        for (int q = 0; q < 4; q++) {
            const uint32_t v = sad16[(4 * q + 0) * n_idx + i] + sad16[(4 * q + 1) * n_idx + i] + sad16[(4 * q + 2) * n_idx + i] +
                               sad16[(4 * q + 3) * n_idx + i];
            sad32[q * n_idx + i] = v;
            if (v < best32[q]) {
                best32[q] = v;
                mv32[q]   = m;
            }
            tot += v;
        }
        if (tot < best64[0]) {
            best64[0] = tot;
            mv64[0]   = m;
        }
    }
}

// one 16x16: 4 threads
__global__ void ext_sad_8_16_kernel(const uint8_t* src, uint32_t ss, const uint8_t* ref, uint32_t rs, uint32_t* best8,
                                    uint32_t* best16, uint32_t* mv8, uint32_t* mv16, uint32_t mv, uint32_t* sad16,
                                    uint32_t* sad8, int sub) {
    __shared__ uint32_t v4[4];
    const int k = threadIdx.x;
    const uint32_t v = sad8x8_global(src + 8 * (k >> 1) * ss + 8 * (k & 1), ss, ref + 8 * (k >> 1) * rs + 8 * (k & 1), rs, sub);
    v4[k]   = v;
    sad8[k] = v;
    if (v < best8[k]) {
        best8[k] = v;
        mv8[k]   = mv;
    }
    __syncthreads();
    if (k == 0) {
        const uint32_t t = v4[0] + v4[1] + v4[2] + v4[3];
        if (t < best16[0]) {
            best16[0] = t;
            mv16[0]   = mv;
        }
        *sad16 = t;
    }
}

// ---------------------------------------------------------------------------------------------
// T2: full-pel search of one 64x64 block over sa_w x sa_h positions, all 85 square PUs at once
// ---------------------------------------------------------------------------------------------
constexpr int kFpThreads = 256;
constexpr int kFpTW = 16, kFpTH = 4;              // positions per chunk
constexpr int kFpLW = (kFpTW + 64 + 3) / 4 + 3;   // words per staged window line (odd-ish pitch)
constexpr int kFpLines = kFpTH + 63;

__device__ __forceinline__ void stage_words(uint32_t* dst, int nwords, const uint8_t* g, int nbytes, int lane, int nlanes) {
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const int       shift = (int)(ga & 3) * 8;
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
    const int       valid = (nbytes + 3) >> 2, last_src = (int)(((ga & 3) + nbytes - 1) >> 2);
    for (int w = lane; w < nwords; w += nlanes) {
        uint32_t v = 0;
        if (w < valid) {
            const uint32_t lo = __ldg(gw + w);
            const uint32_t hi = (shift && (w + 1) <= last_src) ? __ldg(gw + w + 1) : 0u;
            v = __funnelshift_r(lo, hi, shift);
            const int rem = nbytes - (w << 2);
            if (rem < 4) v &= (1u << (rem * 8)) - 1u;
        }
        dst[w] = v;
    }
}

// rows x nwords words of `nbytes`-byte rows at any alignment, all threads of the CTA, one word per step
__device__ __forceinline__ void stage_rows(uint32_t* dst, int nwords, int rows, const uint8_t* g, size_t pitch, int nbytes) {
    const int valid = (nbytes + 3) >> 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < rows * nwords; i += kFpThreads) {
        const int r = i / nwords, w = i - r * nwords;
        const uintptr_t ga = reinterpret_cast<uintptr_t>(g + (size_t)r * pitch);
        const int       shift = (int)(ga & 3) * 8, last_src = (int)(((ga & 3) + nbytes - 1) >> 2);
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
        uint32_t v = 0;
        if (w < valid) {
            const uint32_t lo = __ldg(gw + w);
            const uint32_t hi = (shift && (w + 1) <= last_src) ? __ldg(gw + w + 1) : 0u;
            v = __funnelshift_r(lo, hi, shift);
            const int rem = nbytes - (w << 2);
            if (rem < 4) v &= (1u << (rem * 8)) - 1u;
        }
        dst[r * nwords + w] = v;
    }
}

__global__ void __launch_bounds__(kFpThreads)
fullpel_search_kernel(const uint8_t* __restrict__ src_plane, const uint8_t* __restrict__ ref_plane,
                      const SvtB200FullpelItem* __restrict__ items, int n_items, uint32_t* __restrict__ best_sad,
                      uint32_t* __restrict__ best_mv) {
    __shared__ uint32_t S[64 * 16];
    __shared__ uint32_t W[kFpLines * kFpLW];
    __shared__ uint32_t sad8[kFpTW * kFpTH][65];
    __shared__ uint32_t sadpu[kFpTW * kFpTH][21];  // 64x64, 4 x 32x32, 16 x 16x16 of every position
    __shared__ unsigned long long best[85];

    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const SvtB200FullpelItem item = items[it];
        const int sa_w = item.sa_w, sa_h = item.sa_h, sub = item.sub_sad;
        if (threadIdx.x < 85) best[threadIdx.x] = ((unsigned long long)(128u * 128u * 255u) << 32) | 0xffffffffull;
        // flat staging: every thread's words are independent loads, so one global round trip covers the block
        stage_rows(S, 16, 64, src_plane + item.src_off, item.src_stride, 64);
        for (int y0 = 0; y0 < sa_h; y0 += kFpTH) {
            const int th = min(kFpTH, sa_h - y0);
            for (int x0 = 0; x0 < sa_w; x0 += kFpTW) {
                const int tw = min(kFpTW, sa_w - x0);
                __syncthreads();
                stage_rows(W, kFpLW, th + 63, ref_plane + item.ref_off + (size_t)y0 * item.ref_stride + x0, item.ref_stride, tw + 63);
                __syncthreads();
                // 8x8 SADs: unit = (position, 8x8 block), stored in z-order so that the four 8x8 of a 16x16
                // (and the four 16x16 of a 32x32) are neighbours
                const int npos = tw * th;
                for (int u = threadIdx.x; u < npos * 64; u += kFpThreads) {
                    const int pos = u >> 6, blk = u & 63;
                    const int by = blk >> 3, bx = blk & 7;
                    const int py = pos / tw, px = pos - py * tw;
                    const int a8 = (px & 3) * 8, wb = (px >> 2) + 2 * bx;
                    uint32_t  acc = 0;
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (sub && (r & 1)) continue;
                        const uint32_t* L = W + (py + 8 * by + r) * kFpLW + wb;
                        const uint32_t* Sr = S + (8 * by + r) * 16 + 2 * bx;
                        const uint32_t w0 = L[0], w1 = L[1], w2 = L[2];
                        acc = __vsadu4(Sr[0], __funnelshift_r(w0, w1, a8)) + acc;
                        acc = __vsadu4(Sr[1], __funnelshift_r(w1, w2, a8)) + acc;
                    }
                    if (sub) acc <<= 1;
                    const int y16 = by >> 1, x16 = bx >> 1;
                    sad8[pos][4 * z16_of(y16, x16) + 2 * (by & 1) + (bx & 1)] = acc;
                }
                __syncthreads();
                // 16x16 / 32x32 / 64x64 sums: 16 lanes per position, lane = 16x16 z-index; two shuffle
                // steps give the 32x32 of each quad, two more the 64x64
                for (int u = threadIdx.x; u < ((npos * 16 + 31) & ~31); u += kFpThreads) {
                    const int  pos = u >> 4, z = u & 15;
                    const bool on = pos < npos;
                    uint32_t   v16 = 0;
                    if (on) v16 = sad8[pos][4 * z] + sad8[pos][4 * z + 1] + sad8[pos][4 * z + 2] + sad8[pos][4 * z + 3];
                    uint32_t v32 = v16 + __shfl_xor_sync(0xffffffffu, v16, 1);
                    v32 += __shfl_xor_sync(0xffffffffu, v32, 2);
                    uint32_t v64 = v32 + __shfl_xor_sync(0xffffffffu, v32, 4);
                    v64 += __shfl_xor_sync(0xffffffffu, v64, 8);
                    if (on) {
                        sadpu[pos][5 + z] = v16;
                        if ((z & 3) == 0) sadpu[pos][1 + (z >> 2)] = v32;
                        if (z == 0) sadpu[pos][0] = v64;
                    }
                }
                __syncthreads();
                // per-PU scan of this chunk's positions; key = sad<<32 | raster index
                if (threadIdx.x < 85) {
                    const int pu = threadIdx.x;
                    unsigned long long b = best[pu];
                    for (int pos = 0; pos < npos; pos++) {
                        const uint32_t v = pu < 21 ? sadpu[pos][pu] : sad8[pos][pu - 21];
                        const int py = pos / tw, px = pos - py * tw;
                        const unsigned long long key =
                            ((unsigned long long)v << 32) | (unsigned long long)(uint32_t)((y0 + py) * sa_w + (x0 + px));
                        b = key < b ? key : b;
                    }
                    best[pu] = b;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 85) {
            const unsigned long long b = best[threadIdx.x];
            const uint32_t idx = (uint32_t)(b & 0xffffffffull);
            best_sad[(size_t)it * 85 + threadIdx.x] = (uint32_t)(b >> 32);
            if (idx == 0xffffffffu)
                best_mv[(size_t)it * 85 + threadIdx.x] = 0;
            else {
                const int y = (int)(idx / (uint32_t)sa_w), x = (int)(idx - (uint32_t)y * sa_w);
                best_mv[(size_t)it * 85 + threadIdx.x] = pack_mv(item.org_x + x, item.org_y + y);
            }
        }
        __syncthreads();
    }
}

}  // namespace b200

using namespace b200;

extern "C" void svt_b200_ext_all_sad_calculation_8x8_16x16(uint8_t* src, uint32_t src_stride, uint8_t* ref, uint32_t ref_stride,
                                                           uint32_t mv, uint32_t* p_best_sad_8x8, uint32_t* p_best_sad_16x16,
                                                           uint32_t* p_best_mv8x8, uint32_t* p_best_mv16x16,
                                                           uint32_t p_eight_sad16x16[16][8], uint32_t p_eight_sad8x8[64][8],
                                                           uint8_t sub_sad) {
    (void)p_eight_sad8x8;  // not written by the reference C kernel either (motion_estimation.c:221)
    require_ready();
    LaneGuard l;
    const size_t sb = 63 * (size_t)src_stride + 64, rb = 63 * (size_t)ref_stride + 64 + 7;
    size_t o_src = l->alloc(sb), o_ref = l->alloc(rb);
    size_t o_st = l->alloc((64 + 16 + 64 + 16) * 4);  // best8, best16, mv8, mv16
    size_t in_end = l->used;
    size_t o_e16 = l->alloc(16 * 8 * 4);
    memcpy(l->h<uint8_t>(o_src), src, sb);
    memcpy(l->h<uint8_t>(o_ref), ref, rb);
    uint32_t* st = l->h<uint32_t>(o_st);
    memcpy(st, p_best_sad_8x8, 64 * 4);
    memcpy(st + 64, p_best_sad_16x16, 16 * 4);
    memcpy(st + 80, p_best_mv8x8, 64 * 4);
    memcpy(st + 144, p_best_mv16x16, 16 * 4);
    l->h2d(0, in_end);
    uint32_t* d = l->d<uint32_t>(o_st);
    ext_all_sad_kernel<<<1, 64, 0, l->stream>>>(l->d<uint8_t>(o_src), src_stride, l->d<uint8_t>(o_ref), ref_stride, mv, d, d + 64,
                                                d + 80, d + 144, l->d<uint32_t>(o_e16), sub_sad ? 1 : 0);
    B200_LAUNCH_CHECK();
    l->d2h(o_st, (o_e16 + 16 * 8 * 4) - o_st);
    l->sync();
    memcpy(p_best_sad_8x8, st, 64 * 4);
    memcpy(p_best_sad_16x16, st + 64, 16 * 4);
    memcpy(p_best_mv8x8, st + 80, 64 * 4);
    memcpy(p_best_mv16x16, st + 144, 16 * 4);
    memcpy(p_eight_sad16x16, l->h<uint32_t>(o_e16), 16 * 8 * 4);
}

static void ext_32_64_t1(const uint32_t* sad16, uint32_t* b32, uint32_t* b64, uint32_t* m32, uint32_t* m64, uint32_t mv,
                         uint32_t* sad32, int n_idx) {
    require_ready();
    LaneGuard l;
    size_t o_in = l->alloc(16 * 8 * 4), o_st = l->alloc(10 * 4);
    size_t in_end = l->used;
    size_t o_s32 = l->alloc(4 * 8 * 4);
    memcpy(l->h<uint32_t>(o_in), sad16, (size_t)16 * n_idx * 4);
    uint32_t* st = l->h<uint32_t>(o_st);
    memcpy(st, b32, 16);
    st[4] = b64[0];
    memcpy(st + 5, m32, 16);
    st[9] = m64[0];
    l->h2d(0, in_end);
    uint32_t* d = l->d<uint32_t>(o_st);
    ext_sad_32_64_kernel<<<1, 32, 0, l->stream>>>(l->d<uint32_t>(o_in), d, d + 4, d + 5, d + 9, mv, l->d<uint32_t>(o_s32), n_idx);
    B200_LAUNCH_CHECK();
    l->d2h(o_st, (o_s32 + 4 * 8 * 4) - o_st);
    l->sync();
    memcpy(b32, st, 16);
    b64[0] = st[4];
    memcpy(m32, st + 5, 16);
    m64[0] = st[9];
    memcpy(sad32, l->h<uint32_t>(o_s32), (size_t)4 * n_idx * 4);
}

extern "C" void svt_b200_ext_eight_sad_calculation_32x32_64x64(uint32_t p_sad16x16[16][8], uint32_t* p_best_sad_32x32,
                                                               uint32_t* p_best_sad_64x64, uint32_t* p_best_mv32x32,
                                                               uint32_t* p_best_mv64x64, uint32_t mv, uint32_t p_sad32x32[4][8]) {
    ext_32_64_t1(&p_sad16x16[0][0], p_best_sad_32x32, p_best_sad_64x64, p_best_mv32x32, p_best_mv64x64, mv, &p_sad32x32[0][0], 8);
}
extern "C" void svt_b200_ext_sad_calculation_32x32_64x64(uint32_t* p_sad16x16, uint32_t* p_best_sad_32x32,
                                                         uint32_t* p_best_sad_64x64, uint32_t* p_best_mv32x32,
                                                         uint32_t* p_best_mv64x64, uint32_t mv, uint32_t* p_sad32x32) {
    ext_32_64_t1(p_sad16x16, p_best_sad_32x32, p_best_sad_64x64, p_best_mv32x32, p_best_mv64x64, mv, p_sad32x32, 1);
}

extern "C" void svt_b200_ext_sad_calculation_8x8_16x16(uint8_t* src, uint32_t src_stride, uint8_t* ref, uint32_t ref_stride,
                                                       uint32_t* p_best_sad_8x8, uint32_t* p_best_sad_16x16, uint32_t* p_best_mv8x8,
                                                       uint32_t* p_best_mv16x16, uint32_t mv, uint32_t* p_sad16x16,
                                                       uint32_t* p_sad8x8, uint8_t sub_sad) {
    require_ready();
    LaneGuard l;
    const size_t sb = 15 * (size_t)src_stride + 16, rb = 15 * (size_t)ref_stride + 16;
    size_t o_src = l->alloc(sb), o_ref = l->alloc(rb), o_st = l->alloc(10 * 4);
    size_t in_end = l->used;
    size_t o_out = l->alloc(5 * 4);
    memcpy(l->h<uint8_t>(o_src), src, sb);
    memcpy(l->h<uint8_t>(o_ref), ref, rb);
    uint32_t* st = l->h<uint32_t>(o_st);
    memcpy(st, p_best_sad_8x8, 16);
    st[4] = p_best_sad_16x16[0];
    memcpy(st + 5, p_best_mv8x8, 16);
    st[9] = p_best_mv16x16[0];
    l->h2d(0, in_end);
    uint32_t* d = l->d<uint32_t>(o_st);
    uint32_t* o = l->d<uint32_t>(o_out);
    ext_sad_8_16_kernel<<<1, 4, 0, l->stream>>>(l->d<uint8_t>(o_src), src_stride, l->d<uint8_t>(o_ref), ref_stride, d, d + 4, d + 5,
                                                d + 9, mv, o + 4, o, sub_sad ? 1 : 0);
    B200_LAUNCH_CHECK();
    l->d2h(o_st, (o_out + 20) - o_st);
    l->sync();
    memcpy(p_best_sad_8x8, st, 16);
    p_best_sad_16x16[0] = st[4];
    memcpy(p_best_mv8x8, st + 5, 16);
    p_best_mv16x16[0] = st[9];
    memcpy(p_sad8x8, l->h<uint32_t>(o_out), 16);
    *p_sad16x16 = l->h<uint32_t>(o_out)[4];
}

// svt_initialize_buffer_32bits (aom_dsp_rtcd.h:855): pure host-memory fill; there is nothing for a
// device to do here, so the B200 tier keeps it as the trivial host loop it is.
extern "C" void svt_b200_initialize_buffer_32bits(uint32_t* pointer, uint32_t count128, uint32_t count32, uint32_t value) {
    const uint32_t n = count128 * 4 + count32;
    for (uint32_t i = 0; i < n; i++) pointer[i] = value;
}

extern "C" int svt_b200_fullpel_search_batch_dev(const uint8_t* d_src_plane, const uint8_t* d_ref_plane,
                                                 const SvtB200FullpelItem* d_items, int n_items, uint32_t* d_best_sad,
                                                 uint32_t* d_best_mv, void* stream) {
    require_ready();
    if (n_items <= 0) return n_items == 0 ? SVT_B200_OK : SVT_B200_ERR_BAD_ARG;
    fullpel_search_kernel<<<grid_for(n_items, 4), kFpThreads, 0, (cudaStream_t)stream>>>(d_src_plane, d_ref_plane, d_items, n_items,
                                                                                       d_best_sad, d_best_mv);
    B200_LAUNCH_CHECK();
    return SVT_B200_OK;
}

extern "C" int svt_b200_fullpel_search_batch_host(const uint8_t* src_plane, size_t src_bytes, const uint8_t* ref_plane,
                                                  size_t ref_bytes, const SvtB200FullpelItem* items, int n_items,
                                                  uint32_t* best_sad, uint32_t* best_mv) {
    require_ready();
    if (n_items <= 0) return n_items == 0 ? SVT_B200_OK : SVT_B200_ERR_BAD_ARG;
    LaneGuard l;
    size_t o_src = l->alloc(src_bytes + 8), o_ref = l->alloc(ref_bytes + 8), o_it = l->alloc(sizeof(SvtB200FullpelItem) * n_items);
    size_t in_end = l->used;
    size_t o_sad = l->alloc((size_t)n_items * 85 * 4), o_mv = l->alloc((size_t)n_items * 85 * 4);
    memcpy(l->h<uint8_t>(o_src), src_plane, src_bytes);
    memcpy(l->h<uint8_t>(o_ref), ref_plane, ref_bytes);
    memcpy(l->h<uint8_t>(o_it), items, sizeof(SvtB200FullpelItem) * n_items);
    l->h2d(0, in_end);
    svt_b200_fullpel_search_batch_dev(l->d<uint8_t>(o_src), l->d<uint8_t>(o_ref), l->d<SvtB200FullpelItem>(o_it), n_items,
                                      l->d<uint32_t>(o_sad), l->d<uint32_t>(o_mv), l->stream);
    l->d2h(o_sad, (o_mv + (size_t)n_items * 85 * 4) - o_sad);
    l->sync();
    memcpy(best_sad, l->h<uint32_t>(o_sad), (size_t)n_items * 85 * 4);
    memcpy(best_mv, l->h<uint32_t>(o_mv), (size_t)n_items * 85 * 4);
    return SVT_B200_OK;
}
