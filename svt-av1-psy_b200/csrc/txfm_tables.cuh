// txfm_tables.cuh -- constant tables + the uniform 1-D butterfly evaluator shared by the forward
// (txfm_fwd.cu) and inverse (txfm_inv.cu) 2-D transform kernels.
//
// AV1's 1-D DCT/ADST are multi-stage butterfly networks with a rounding shift after every rotation
// (reference: half_btf, Source/Lib/Codec/inv_transforms.h:264), so a GEMM cannot reproduce them
// bit-exactly.  We therefore run the network itself, but not as straight-line scalar
// code per vector: the network is DATA (txfm_graphs.inc: one packed node per (stage, element)) and a
// whole team of threads evaluates one stage of ALL columns (or rows) of a block at once out of
// shared memory -- element-major layout, vector index fastest, so every access is conflict-free and
// every lane runs the same instruction stream (add/sub are evaluated as weight +-1 rotations with a
// zero rounding shift, so there is no divergence between butterfly kinds).
#pragma once
#include <cstdint>
#include "common.cuh"

namespace b200 {

enum {  // TxfmType, Source/Lib/Codec/inv_transforms.h:91-108
    TT_DCT4, TT_DCT8, TT_DCT16, TT_DCT32, TT_DCT64, TT_ADST4, TT_ADST8, TT_ADST16, TT_ADST32,
    TT_IDTX4, TT_IDTX8, TT_IDTX16, TT_IDTX32, TT_IDTX64, TT_TYPES
};

struct GraphDesc {
    int off;     // first node in g_txfm_nodes
    int n;       // points
    int stages;  // 0 => not a butterfly network (ADST4 / identity: closed forms below)
};

struct TxCfg {  // per (tx_size, tx_type); dumped from the reference by tools/dump_txfm_cfg.py
    int8_t valid;
    int8_t f_ud, f_lr, f_s0, f_s1, f_s2, f_cbc, f_cbr, f_tc, f_tr;
    int8_t i_ud, i_lr, i_s0, i_s1, i_cbc, i_cbr, i_tc, i_tr;
};

// NOTE: this header is included by exactly one translation unit (txfm.cu): the symbols below are
// defined here so that no relocatable device code is needed.
__constant__ GraphDesc c_graph[2][TT_TYPES];  // [0]=forward, [1]=inverse
__constant__ TxCfg     c_txcfg[19][16];
__constant__ int32_t   c_cospi[7][64];  // bit 10..16
__constant__ int32_t   c_sinpi[7][5];

// packed node: a[0:6) b[6:12) wa[12:20) wb[20:28) is_rotation[28] clamp[29]
#define TXG_BEGIN(tag, n, st)
#define TXG_END(tag)
#define TXG_NODE(btf, wa, a, wb, b, cl)                                                                      \
    ((uint32_t)(a) | ((uint32_t)(b) << 6) | (((uint32_t)(wa) & 0xffu) << 12) | (((uint32_t)(wb) & 0xffu) << 20) | \
     ((uint32_t)(btf) << 28) | ((uint32_t)(cl) << 29)),
__device__ const uint32_t g_txfm_nodes[] = {
#include "txfm_graphs.inc"
    0u};
#undef TXG_BEGIN
#undef TXG_END
#undef TXG_NODE

__host__ __device__ constexpr int tx_w(int s) {
    constexpr int W[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
    return W[s];
}
__host__ __device__ constexpr int tx_h(int s) {
    constexpr int H[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
    return H[s];
}
const TxCfg& host_txcfg(int size, int type);

constexpr int kNewSqrt2    = 5793;  // 2^12 * sqrt(2)
constexpr int kNewInvSqrt2 = 2896;  // 2^12 / sqrt(2)

__device__ __forceinline__ int32_t round_shift64(long long v, int bit) { return (int32_t)((v + (1ll << (bit - 1))) >> bit); }
// svt_av1_round_shift_array_c semantics (inv_transforms.c:2421): bit>0 rounding right shift, bit<0 multiply
__device__ __forceinline__ int32_t round_shift_arr(int32_t v, int bit) {
    if (bit == 0) return v;
    if (bit > 0) return round_shift64((long long)v, bit);
    return (int32_t)((uint32_t)v * (1u << (-bit)));
}
__device__ __forceinline__ int32_t clamp_bits(int32_t v, int bit) {
    if (bit <= 0) return v;
    const int32_t hi = (int32_t)((1ll << (bit - 1)) - 1), lo = (int32_t)(-(1ll << (bit - 1)));
    return v < lo ? lo : (v > hi ? hi : v);
}

template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if constexpr (TEAM == 32)
        __syncwarp();
    else
        __syncthreads();
}

// One 1-D pass over V vectors of N points held element-major in shared memory (x[i*P + v]).
// Returns the buffer holding the result (x or y).  `inverse` selects the network family;
// `clampb` is the inverse path's per-stage clamp width (svt_av1_gen_inv_stage_range,
// inv_transforms.c:42-83: the same width for every stage of a pass).
template <int TEAM>
__device__ int32_t* txfm_pass_1d(int type, int inverse, int32_t* x, int32_t* y, int N, int V, int P, int cos_bit,
                                 int clampb, int tid) {
    const GraphDesc g    = c_graph[inverse][type];
    const int       tot  = N * V;
    const int       lgV  = 31 - __clz(V);  // V is a power of two (4..64)
    if (g.stages > 0) {
        const int32_t* cosv = c_cospi[cos_bit - 10];
        const long long rnd = 1ll << (cos_bit - 1);
        for (int s = 0; s < g.stages; s++) {
            const uint32_t* nodes = g_txfm_nodes + g.off + s * N;
            for (int idx = tid; idx < tot; idx += TEAM) {
                const int      i  = idx >> lgV, v = idx & (V - 1);
                const uint32_t nd = __ldg(nodes + i);
                const int      a = nd & 63, b = (nd >> 6) & 63;
                const int      wa = (int)(int8_t)(nd >> 12), wb = (int)(int8_t)(nd >> 20);
                const int32_t  xa = x[a * P + v], xb = x[b * P + v];
                int32_t        r;
                if (nd & (1u << 28)) {  // rotation: half_btf with 32-bit products, 64-bit sum
                    int32_t ca = cosv[wa < 0 ? -wa : wa];
                    int32_t cb = cosv[wb < 0 ? -wb : wb];
                    ca         = wa < 0 ? -ca : ca;
                    cb         = wb < 0 ? -cb : cb;
                    const int32_t p0 = (int32_t)((uint32_t)ca * (uint32_t)xa);
                    const int32_t p1 = (int32_t)((uint32_t)cb * (uint32_t)xb);
                    r                = (int32_t)(((long long)p0 + (long long)p1 + rnd) >> cos_bit);
                } else {
                    r = (int32_t)((uint32_t)(wa * xa) + (uint32_t)(wb * xb));
                    if (nd & (1u << 29)) r = clamp_bits(r, clampb);
                }
                y[i * P + v] = r;
            }
            team_sync<TEAM>();
            int32_t* t = x;
            x          = y;
            y          = t;
        }
        return x;
    }
    if (type == TT_ADST4) {
        // Closed form of svt_av1_fadst4_new (transforms.c:1415-1502) / svt_av1_iadst4_new
        // (inv_transforms.c:722-806): every intermediate is a 32-bit wrapping linear combination,
        // so the four outputs are evaluated directly (mod 2^32) and rounded once.
        const int32_t* sp = c_sinpi[cos_bit - 10];
        const uint32_t s1 = sp[1], s2 = sp[2], s3 = sp[3], s4 = sp[4];
        for (int idx = tid; idx < tot; idx += TEAM) {
            const int      i = idx >> lgV, v = idx & (V - 1);
            const uint32_t x0 = x[v], x1 = x[P + v], x2 = x[2 * P + v], x3 = x[3 * P + v];
            uint32_t       o;
            if (!inverse) {
                if (i == 0) o = s1 * x0 + s2 * x1 + s3 * x2 + s4 * x3;
                else if (i == 1) o = s3 * (x0 + x1 - x3);
                else if (i == 2) o = s4 * x0 - s1 * x1 - s3 * x2 + s2 * x3;
                else o = (s4 * x0 - s1 * x1 + s2 * x3) - (s1 * x0 + s2 * x1 + s4 * x3) + s3 * x2;
            } else {
                // a = s1*x0 + s4*x2 + s2*x3 ; b = s2*x0 - s1*x2 - s4*x3 ; c = s3*x1 ; d = s3*(x0 - x2 + x3)
                const uint32_t a = s1 * x0 + s4 * x2 + s2 * x3;
                const uint32_t b = s2 * x0 - s1 * x2 - s4 * x3;
                const uint32_t c = s3 * x1;
                if (i == 0) o = a + c;
                else if (i == 1) o = b + c;
                else if (i == 2) o = s3 * (x0 - x2 + x3);
                else o = a + b - c;
            }
            y[i * P + v] = round_shift64((long long)(int32_t)o, cos_bit);
        }
        team_sync<TEAM>();
        return y;
    }
    // identity kernels (transforms.c:2205-2236, inv_transforms.c:2331-2362)
    for (int idx = tid; idx < tot; idx += TEAM) {
        const int     i = idx >> lgV, v = idx & (V - 1);
        const int32_t xv = x[i * P + v];
        int32_t       r;
        switch (type) {
        case TT_IDTX4: r = round_shift64((long long)xv * kNewSqrt2, 12); break;
        case TT_IDTX8: r = (int32_t)((uint32_t)xv * 2u); break;
        case TT_IDTX16: r = round_shift64((long long)xv * 2 * kNewSqrt2, 12); break;
        case TT_IDTX32: r = (int32_t)((uint32_t)xv * 4u); break;
        default: r = round_shift64((long long)xv * 4 * kNewSqrt2, 12); break;
        }
        y[i * P + v] = r;
    }
    team_sync<TEAM>();
    return y;
}

}  // namespace b200
