#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const CUtensorMap* tm, uint8_t* out, int x, int y, int z, int variant) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* tile = sm + ((128u - (smem_u32(sm) & 127u)) & 127u);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(256 * 33) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 256 * 33; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int pitch = 704, rows = 518, F = 2;
    std::vector<uint8_t> h((size_t)pitch * rows * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / pitch);
    uint8_t *d, *dout; cudaMalloc(&d, h.size()); cudaMalloc(&dout, 256 * 33);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry: %d %p %d\n", (int)e, fn, (int)q);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)F};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    cuuint32_t box[3] = {256, 33, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    int tests[4][3] = {{0, 0, 0}, {16, 3, 1}, {45, 35, 1}, {500, 500, 0}};
    for (auto& t : tests) {
        k<<<1, 256, 256 * 33 + 256>>>(dtm, dout, t[0], t[1], t[2], 0);
        cudaError_t err = cudaDeviceSynchronize();
        std::vector<uint8_t> o(256 * 33);
        cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < 33; ++yy) for (int xx = 0; xx < 256; ++xx) {
            int gx = t[0] + xx, gy = t[1] + yy;
            uint8_t want = (gx < pitch && gy < rows) ? h[(size_t)t[2] * pitch * rows + (size_t)gy * pitch + gx] : 0;
            bad += o[yy * 256 + xx] != want;
        }
        printf("x=%d y=%d z=%d: err=%s bad=%d\n", t[0], t[1], t[2], cudaGetErrorString(err), bad);
    }
    return 0;
}
