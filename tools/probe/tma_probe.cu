// tools/probe/tma_probe.cu -- stand-alone probe of cp.async.bulk.tensor.2d on this GPU: one box load per run,
// parameters from the command line:  tma_probe <box_w> <box_h> <c0> <c1> <pitch> <rows> <map_in: 0 param | 1 global>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Maps { CUtensorMap m; };
__global__ void probe(const __grid_constant__ Maps maps, const CUtensorMap* gmap, int use_global, int c0, int c1, int bytes, uint32_t* out) {
    __shared__ __align__(128) uint8_t buf[256 * 80];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        const CUtensorMap* mp = use_global ? gmap : &maps.m;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(mp)), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) { out[0] = buf[0]; out[1] = buf[1]; out[2] = buf[bytes - 1]; }
}
int main(int argc, char** argv) {
    int bw = atoi(argv[1]), bh = atoi(argv[2]), c0 = atoi(argv[3]), c1 = atoi(argv[4]), pitch = atoi(argv[5]), rows = atoi(argv[6]), g = atoi(argv[7]);
    uint8_t* d; cudaMalloc(&d, (size_t)pitch * rows);
    uint8_t* h = (uint8_t*)malloc((size_t)pitch * rows);
    for (size_t i = 0; i < (size_t)pitch * rows; i++) h[i] = (uint8_t)(i * 7 + (i / pitch));
    cudaMemcpy(d, h, (size_t)pitch * rows, cudaMemcpyHostToDevice);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Maps maps; memset(&maps, 0, sizeof(maps));
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CUresult r = ((Fn)fnp)(&maps.m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d ", (int)r, (int)q);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &maps.m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    uint32_t* out; cudaMalloc(&out, 16);
    probe<<<1, 32>>>(maps, gm, g, c0, c1, bw * bh, out);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t ho[3] = {0, 0, 0}; cudaMemcpy(ho, out, 12, cudaMemcpyDeviceToHost);
    printf("box %dx%d at (%d,%d) pitch %d map_in=%s -> %s  got %u %u want %u %u\n", bw, bh, c0, c1, pitch, g ? "global" : "param", cudaGetErrorString(e), ho[0], ho[1],
           h[(size_t)c1 * pitch + c0], h[(size_t)c1 * pitch + c0 + 1]);
    return 0;
}
