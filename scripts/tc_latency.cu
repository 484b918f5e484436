// tc_latency.cu -- how long does "issue n tcgen05.mma + commit + mbarrier wait" take? (GPU box only)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I situation3d_b200/csrc -o scripts/tc_latency scripts/tc_latency.cu
#include "tc_common.cuh"
#include <cstdio>
using namespace pn2;

// mode 0: all threads wait on the mbarrier; mode 1: only thread 0 waits, the rest sit in bar.sync
__global__ void __launch_bounds__(256) k(int nmma, int N, int mode, int nthreads_wait, long long *out)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;
    if (tid == 0) { tc_mbar_init(smem_u32(&mbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<256>(smem_u32(&slot));
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot, a = smem_u32(base), b = smem_u32(base) + 32768;
    const uint32_t idesc = umma_idesc(128, N);
    uint32_t phase = 0;
    long long tot = 0, tissue = 0;
    for (int it = 0; it < 200; ++it) {
        __syncthreads();
        const long long t0 = clock64();
        if (mode < 2) {
            if (tid == 0) {
                tc_fence_after();
                for (int ks = 0; ks < nmma; ++ks)
                    umma_bf16(tmem, smem_desc(a + (ks & 3) * 32 + (ks >> 2) * 16384, 1024u, kSw128),
                              smem_desc(b + (ks & 3) * 32 + (ks >> 2) * 16384, 1024u, kSw128), idesc, ks > 0);
                umma_commit(smem_u32(&mbar));
            }
        } else if (warp == 0) {
            // warp-uniform issue: descriptors are computed by the whole warp, one elected lane issues
            tc_fence_after();
            uint32_t elected;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(elected));
            const uint64_t da0 = smem_desc(a, 1024u, kSw128), db0 = smem_desc(b, 1024u, kSw128);
            for (int ks = 0; ks < nmma; ++ks) {
                const uint64_t off = (uint64_t)(((ks & 3) * 32 + (ks >> 2) * 16384) >> 4);
                if (elected) umma_bf16(tmem, da0 + off, db0 + off, idesc, ks > 0);
            }
            if (elected) umma_commit(smem_u32(&mbar));
            __syncwarp();
        }
        const long long t1 = clock64();
        if (mode != 1 || tid == 0) tc_mbar_wait(smem_u32(&mbar), phase);
        phase ^= 1;
        tc_fence_after();
        const long long t2 = clock64();
        if (it >= 20) { tot += t2 - t0; tissue += t1 - t0; }
        tc_fence_before();
    }
    if (tid == 0) { out[0] = tot / 180; out[1] = tissue / 180; }
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

int main()
{
    long long *out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    printf("threads mode N nmma : cycles(issue+commit+wait) issue-only\n");
    for (int threads : {128, 256})
        for (int mode : {0, 2})
            for (int N : {64, 128})
                for (int n : {1, 4, 9, 16}) {
                    k<<<1, threads, 100 * 1024>>>(n, N, mode, threads, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("%4d %d %3d %2d : %6lld %6lld\n", threads, mode, N, n, out[0], out[1]);
                }
    return 0;
}
