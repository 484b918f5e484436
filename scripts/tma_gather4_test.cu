// tma_gather4_test.cu -- what box shape does cp.async.bulk.tensor.2d...tile::gather4 want, and where do the rows land? (GPU box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tma_gather4_test scripts/tma_gather4_test.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, uint32_t *out)
{
    __shared__ __align__(1024) unsigned char buf[1024];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 256; ++i) reinterpret_cast<uint32_t *>(buf)[i] = 0xdeadbeefu;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(512u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                     " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(d), "l"(&tm), "r"(b), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        uint32_t done = 0;
        long long t0 = clock64();
        while (!done && clock64() - t0 < 200000000LL) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(b), "r"(0u) : "memory");
        }
        out[256] = done;
        for (int i = 0; i < 256; ++i) out[i] = reinterpret_cast<uint32_t *>(buf)[i];
    }
}

int main()
{
    const int rows = 1000, cols = 64;
    std::vector<__nv_bfloat16> h(rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) h[r * cols + c] = __float2bfloat16((float)(r % 256) + c / 64.0f);
    __nv_bfloat16 *dt; uint32_t *dout;
    cudaMalloc(&dt, h.size() * 2); cudaMalloc(&dout, 257 * 4);
    cudaMemcpy(dt, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    for (int boxrows : {1, 4}) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)cols * 2};
        cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)boxrows}, estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode rc=%d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(dout, 0, 257 * 4);
        k<<<1, 32>>>(tm, 5, 900, 17, 333, dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        uint32_t o[257];
        cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
        printf("  barrier completed: %u\n", o[256]);
        // print, for each 128-byte smem row and each 16-byte chunk, the source row number of its first element
        for (int sr = 0; sr < 8; ++sr) {
            printf("  smem row %d:", sr);
            for (int ch = 0; ch < 8; ++ch) {
                const uint32_t w = o[sr * 32 + ch * 4];
                if (w == 0xdeadbeefu) { printf("   ----"); continue; }
                __nv_bfloat16 v; *reinterpret_cast<uint16_t *>(&v) = (uint16_t)(w & 0xffff);
                printf(" %6.3f", __bfloat162float(v));
            }
            printf("\n");
        }
    }
    return 0;
}
