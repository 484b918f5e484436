// tmem_bw.cu -- TMEM read / write throughput of tcgen05.ld / tcgen05.st as the fused SA kernel uses them (GPU box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I situation3d_b200/csrc -o scripts/tmem_bw scripts/tmem_bw.cu
// Each of W warps (warp w reads lane quarter w % 4) issues `reps` x tcgen05.ld.32x32b.x32 (4 KB each) back to back.
// Reports SM cycles per instruction and bytes per cycle per SM, alone and while another warp issues MMAs.
#include "tc_common.cuh"
#include <cstdio>
using namespace pn2;

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32])
{
    tmem_ld32_issue(taddr, v);
}
__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// mode 0: ld x32 ; mode 1: ld x64 ; mode 2: st x32 ; mma != 0: warp 15 issues N=128 K=16 MMAs in a loop meanwhile
__global__ void __launch_bounds__(512) k(int nwarps, int mode, int mma, int reps, long long *out, uint32_t *sink)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;
    if (tid == 0) { tc_mbar_init(smem_u32(&mbar), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t my = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
        uint32_t v[64];
        for (int j = 0; j < 64; ++j) v[j] = j + tid;
        __syncwarp();
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (mode == 0) {
                tmem_ld32_issue(my + (r & 1) * 32, *reinterpret_cast<uint32_t (*)[32]>(&v[0]));
                tmem_ld_wait();
                acc += v[r & 31];
            } else if (mode == 1) {
                tmem_ld64_issue(my, v);
                tmem_ld_wait();
                acc += v[r & 63];
            } else {
                tmem_st32(my + (r & 1) * 32, *reinterpret_cast<uint32_t (*)[32]>(&v[0]));
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        t1 = clock64();
        if (warp == 0 && lane == 0) stop = 1;
    } else if (warp == 15 && mma) {
        const uint32_t idesc = umma_idesc(128, 128);
        const uint32_t a = smem_u32(base), b = smem_u32(base) + 32768;
        uint32_t phase = 0;
        const uint32_t elected = elect_one();
        long long n = 0;
        while (!stop) {
            tc_fence_after();
            for (int ks = 0; ks < 8; ++ks)
                if (elected) umma_bf16(tmem + 256, smem_desc(a + (ks & 3) * 32, 1024u, kSw128), smem_desc(b + (ks & 3) * 32, 1024u, kSw128), idesc, ks > 0);
            if (elected) umma_commit(smem_u32(&mbar));
            __syncwarp();
            tc_mbar_wait(smem_u32(&mbar), phase);
            phase ^= 1;
            n += 8;
        }
        if (lane == 0) out[17] = n;
    }
    if (lane == 0 && warp < 16) out[warp] = t1 - t0;
    sink[tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main()
{
    long long *out; uint32_t *sink;
    cudaMallocManaged(&out, 64 * 8);
    cudaMalloc(&sink, 512 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 2000;
    printf("mode(0=ld.x32 1=ld.x64 2=st.x32) warps mma : cycles/instr (slowest warp)  bytes/cycle/SM   [mma issued]\n");
    for (int mode : {0, 1, 2})
        for (int mma : {0, 1})
            for (int w : {1, 2, 4, 8}) {
                for (int i = 0; i < 20; ++i) out[i] = 0;
                k<<<1, 512, 100 * 1024>>>(w, mode, mma, reps, out, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long mx = 0;
                for (int i = 0; i < w; ++i) mx = out[i] > mx ? out[i] : mx;
                const double per = (double)mx / reps;
                const double bytes = (mode == 1 ? 8192.0 : 4096.0) * w;
                printf("%d %d %d : %8.1f  %8.1f   [%lld]\n", mode, w, mma, per, bytes / per, mma ? out[17] : 0LL);
            }
    return 0;
}
