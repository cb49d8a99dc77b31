// wiener_unit.cuh -- the arithmetic of one separable Wiener unit (<= 64 x 64) out of a staged shared-memory tile:
// shared by wiener_convolve_kernel (wiener.cu: plain units) and lr_filter_kernel (restoration.cu: stripe-aware units).
// s_src: rows -3 .. h+3, columns -3 .. w+4 of the input at pitch 72 (uint16); s_tmp: (64 + 8) x 64 uint16 scratch.
// Follows svt_av1_wiener_convolve_add_src_c / svt_av1_highbd_wiener_convolve_add_src_c (convolve.c:100-147, 194-237).
#pragma once
#include <cstdint>

namespace b200 {

__device__ __forceinline__ int round_pow2_s(int v, int n) { return (v + ((1 << n) >> 1)) >> n; }

template <typename PIX>
__device__ __forceinline__ void wiener_unit_compute(const uint16_t* s_src, uint16_t* s_tmp, PIX* dst, int dst_stride, int w, int h,
                                                    const int16_t* hfilter, const int16_t* vfilter, int bd, int round0, int round1,
                                                    int lbd_rows) {
    const int sh = h + 7;
    const int limit = (1 << (bd + 1 + 7 - round0)) - 1;
    // horizontal: intermediate rows -3..h+3; the low-bit-depth reference computes h+6 rows and
    // zero-fills the last one (convolve.c:113-121)
    const int hrows = lbd_rows ? h + 6 : h + 7;
    const int pmax = (1 << bd) - 1;
    if ((w & 3) == 0) {
        // four adjacent outputs per thread: 3 (horizontal) / 8 (vertical) 64-bit shared loads feed 32 MACs
        int hf[8], vf[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            hf[k] = hfilter[k];
            vf[k] = vfilter[k];
        }
        const int w4 = w >> 2;
        for (int i = threadIdx.x; i < sh * w4; i += blockDim.x) {
            const int r = i / w4, c = (i - r * w4) * 4;
            uint2 out = make_uint2(0u, 0u);
            if (r < hrows) {
                const uint2* p = reinterpret_cast<const uint2*>(s_src + r * 72 + c);
                const uint2  a = p[0], b = p[1], d = p[2];
                const int    x[12] = {(int)(a.x & 0xffff), (int)(a.x >> 16), (int)(a.y & 0xffff), (int)(a.y >> 16),
                                      (int)(b.x & 0xffff), (int)(b.x >> 16), (int)(b.y & 0xffff), (int)(b.y >> 16),
                                      (int)(d.x & 0xffff), (int)(d.x >> 16), (int)(d.y & 0xffff), (int)(d.y >> 16)};
                int v[4];
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    int sum = (x[o + 3] << 7) + (1 << (bd + 6));
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += x[o + k] * hf[k];
                    const int t = round_pow2_s(sum, round0);
                    v[o] = t < 0 ? 0 : (t > limit ? limit : t);
                }
                out = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
            }
            *reinterpret_cast<uint2*>(s_tmp + r * 64 + c) = out;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < h * w4; i += blockDim.x) {
            const int r = i / w4, c = (i - r * w4) * 4;
            const uint2 mid = *reinterpret_cast<const uint2*>(s_tmp + (r + 3) * 64 + c);
            int sum[4] = {(int)(mid.x & 0xffff) << 7, (int)(mid.x >> 16) << 7, (int)(mid.y & 0xffff) << 7, (int)(mid.y >> 16) << 7};
#pragma unroll
            for (int k = 0; k < 8; k++) {  // r + k <= h + 6 < sh: always a staged row
                const uint2 t = *reinterpret_cast<const uint2*>(s_tmp + (r + k) * 64 + c);
                sum[0] += (int)(t.x & 0xffff) * vf[k];
                sum[1] += (int)(t.x >> 16) * vf[k];
                sum[2] += (int)(t.y & 0xffff) * vf[k];
                sum[3] += (int)(t.y >> 16) * vf[k];
            }
#pragma unroll
            for (int o = 0; o < 4; o++) {
                int v = round_pow2_s(sum[o] - (1 << (bd + round1 - 1)), round1);
                v = v < 0 ? 0 : (v > pmax ? pmax : v);
                dst[(ptrdiff_t)r * dst_stride + c + o] = (PIX)v;
            }
        }
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < sh * w; i += blockDim.x) {
        const int r = i / w, c = i - r * w;
        int v = 0;
        if (r < hrows) {
            int sum = ((int)s_src[r * 72 + c + 3] << 7) + (1 << (bd + 6));
#pragma unroll
            for (int k = 0; k < 8; k++) sum += (int)s_src[r * 72 + c + k] * (int)hfilter[k];
            v = round_pow2_s(sum, round0);
            v = v < 0 ? 0 : (v > limit ? limit : v);
        }
        s_tmp[r * 64 + c] = (uint16_t)v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
        const int r = i / w, c = i - r * w;
        int sum = ((int)s_tmp[(r + 3) * 64 + c] << 7) - (1 << (bd + round1 - 1));
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (r + k < sh) sum += (int)s_tmp[(r + k) * 64 + c] * (int)vfilter[k];
        int v = round_pow2_s(sum, round1);
        v = v < 0 ? 0 : (v > pmax ? pmax : v);
        dst[(ptrdiff_t)r * dst_stride + c] = (PIX)v;
    }
    __syncthreads();
}

}  // namespace b200
