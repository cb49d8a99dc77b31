// sgr_unit.cuh -- the self-guided filter of one processing unit out of a staged shared-memory tile (tables + sgr_pass):
// shared by sgr_filter_kernel (sgr.cu) and lr_filter_kernel (restoration.cu).  See sgr.cu for the reference citations.
#pragma once
#include <cstdint>
#include "sgr_tables.inc"

namespace b200 {

static __constant__ int c_sgr_params[64] = SGR_PARAMS_INIT;
static __constant__ int c_x_by_xplus1[256] = SGR_X_BY_XPLUS1_INIT;
static __constant__ int c_one_by_x[25] = SGR_ONE_BY_X_INIT;
static const int h_sgr_params[64] = SGR_PARAMS_INIT;

__device__ __forceinline__ uint32_t rp2u(uint32_t v, int n) { return n ? (v + ((1u << n) >> 1)) >> n : v; }
__device__ __forceinline__ int32_t rp2s(int32_t v, int n) { return (v + ((1 << n) >> 1)) >> n; }

// tile: (h+6) x (w+6) pixels as uint16, pitch tp; A/B: (h+2) x (w+2) int32, pitch ap, origin = (-1,-1)
static __device__ void sgr_pass(const uint16_t* tile, int tp, int w, int h, int32_t* A, int32_t* B, int ap, int r, uint32_t s, int bd,
                         int fast, int32_t* dst, int dst_stride) {
    const int n = (2 * r + 1) * (2 * r + 1);
    const int aw = w + 2, ah = h + 2;
    for (int idx = threadIdx.x; idx < aw * ah; idx += blockDim.x) {
        const int ii = idx / aw, jj = idx - ii * aw;  // ii = i + 1, jj = j + 1
        if (fast && (ii & 1)) continue;               // i = ii - 1 must be odd -> ii even
        uint32_t sum = 0, sq = 0;
        const uint16_t* c = tile + (ii + 2) * tp + (jj + 2);  // pixel (i, j) sits at tile (i+3, j+3)
        for (int dy = -r; dy <= r; dy++)
            for (int dx = -r; dx <= r; dx++) {
                const uint32_t v = c[dy * tp + dx];
                sum += v;
                sq += v * v;
            }
        const uint32_t a = rp2u(sq, 2 * (bd - 8)), b = rp2u(sum, bd - 8);
        const uint32_t p = (a * n < b * b) ? 0u : a * n - b * b;
        const uint32_t z = rp2u(p * s, 20);
        const int      Av = c_x_by_xplus1[z < 255u ? z : 255u];
        A[ii * ap + jj] = Av;
        B[ii * ap + jj] = (int32_t)rp2u((uint32_t)(256 - Av) * sum * (uint32_t)c_one_by_x[n - 1], 12);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < w * h; idx += blockDim.x) {
        const int i = idx / w, j = idx - i * w;
        const int32_t* a0 = A + (i + 1) * ap + (j + 1);
        const int32_t* b0 = B + (i + 1) * ap + (j + 1);
        int32_t a, b, nb;
        if (fast) {
            if (!(i & 1)) {
                nb = 5;
                a = (a0[-ap] + a0[ap]) * 6 + (a0[-1 - ap] + a0[-1 + ap] + a0[1 - ap] + a0[1 + ap]) * 5;
                b = (b0[-ap] + b0[ap]) * 6 + (b0[-1 - ap] + b0[-1 + ap] + b0[1 - ap] + b0[1 + ap]) * 5;
            } else {
                nb = 4;
                a = a0[0] * 6 + (a0[-1] + a0[1]) * 5;
                b = b0[0] * 6 + (b0[-1] + b0[1]) * 5;
            }
        } else {
            nb = 5;
            a = (a0[0] + a0[-1] + a0[1] + a0[-ap] + a0[ap]) * 4 + (a0[-1 - ap] + a0[-1 + ap] + a0[1 - ap] + a0[1 + ap]) * 3;
            b = (b0[0] + b0[-1] + b0[1] + b0[-ap] + b0[ap]) * 4 + (b0[-1 - ap] + b0[-1 + ap] + b0[1 - ap] + b0[1 + ap]) * 3;
        }
        const int32_t v = a * (int32_t)tile[(i + 3) * tp + (j + 3)] + b;
        dst[(size_t)i * dst_stride + j] = rp2s(v, 8 + nb - 4);
    }
    __syncthreads();
}

}  // namespace b200
