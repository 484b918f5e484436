// ubench.cu -- latency micro-benchmarks that size the FPS round (GPU box only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 2000

__global__ void k_redux(uint32_t *out, long long *cyc) {
    uint32_t v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) v = __reduce_max_sync(0xffffffffu, v + i) ^ threadIdx.x;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[0] = (t1 - t0) / N_IT; }
}
__global__ void k_shfl(uint32_t *out, long long *cyc) {
    uint32_t v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) v = max(v, __shfl_xor_sync(0xffffffffu, v, (i & 15) + 1)) + 1;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[1] = (t1 - t0) / N_IT; }
}
__global__ void k_ballot(uint32_t *out, long long *cyc) {
    uint32_t v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) v = __ballot_sync(0xffffffffu, (v >> (i & 7)) & 1) + threadIdx.x;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[2] = (t1 - t0) / N_IT; }
}
__global__ void k_bar(uint32_t *out, long long *cyc, int slot) {
    __shared__ uint32_t s[32];
    uint32_t v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) {
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v + i;
        __syncthreads();
        v += s[(threadIdx.x + i) & ((blockDim.x >> 5) - 1)];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[slot] = (t1 - t0) / N_IT; }
}
__global__ void k_udiv(uint32_t *out, long long *cyc, uint32_t d) {
    uint32_t v = threadIdx.x * 2654435761u + 12345;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) v = v / d + v * 3 + i;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[6] = (t1 - t0) / N_IT; }
}
// argmax round as in fps: compute-free, 64-bit key via two redux + ballot + smem stage
__global__ void k_round2(uint32_t *out, long long *cyc, int slot) {
    __shared__ uint2 wk[2][32];
    __shared__ float4 wx[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    uint32_t hi = threadIdx.x * 2654435761u, lo = ~threadIdx.x;
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) {
        const int buf = i & 1;
        uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
        uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        if (hi == mh && lo == ml) { wk[buf][warp] = make_uint2(mh, ml); wx[buf][warp] = make_float4(acc, 1.f, 2.f, 0.f); }
        __syncthreads();
        const uint2 k = lane < W ? wk[buf][lane] : make_uint2(0u, 0u);
        mh = __reduce_max_sync(0xffffffffu, k.x);
        ml = __reduce_max_sync(0xffffffffu, k.x == mh ? k.y : 0u);
        const int src = __ffs(__ballot_sync(0xffffffffu, k.x == mh && k.y == ml)) - 1;
        const float4 v = wx[buf][src];
        acc += v.x + v.y;
        hi = hi * 1664525u + __float_as_uint(acc);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = hi; cyc[slot] = (t1 - t0) / N_IT; }
}
// same with one redux + ballot (position-ordered tie-break) and a 16-byte message
__global__ void k_round1(uint32_t *out, long long *cyc, int slot) {
    __shared__ float4 wx[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    uint32_t hi = threadIdx.x * 2654435761u;
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) {
        const int buf = i & 1;
        uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
        int src = __ffs(__ballot_sync(0xffffffffu, hi == mh)) - 1;
        if (lane == src) wx[buf][warp] = make_float4(__uint_as_float(mh), acc, 1.f, 2.f);
        __syncthreads();
        const float4 k = lane < W ? wx[buf][lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        mh = __reduce_max_sync(0xffffffffu, __float_as_uint(k.x));
        src = __ffs(__ballot_sync(0xffffffffu, __float_as_uint(k.x) == mh)) - 1;
        acc += __shfl_sync(0xffffffffu, k.y, src) + __shfl_sync(0xffffffffu, k.z, src);
        hi = hi * 1664525u + __float_as_uint(acc);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = hi; cyc[slot] = (t1 - t0) / N_IT; }
}
// butterfly shuffle argmax (value + index), 5 steps
__global__ void k_bfly(uint32_t *out, long long *cyc) {
    uint32_t v = threadIdx.x * 2654435761u, ix = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N_IT; ++i) {
        uint32_t a = v, b = ix;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
            if (a2 > a || (a2 == a && b2 < b)) { a = a2; b = b2; }
        }
        v = v * 1664525u + a + b;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[7] = (t1 - t0) / N_IT; }
}

int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 64); cudaMallocManaged(&cyc, 32 * sizeof(long long));
    for (int i = 0; i < 32; ++i) cyc[i] = -1;
    k_redux<<<1, 32>>>(out, cyc); k_shfl<<<1, 32>>>(out, cyc); k_ballot<<<1, 32>>>(out, cyc);
    k_bar<<<1, 128>>>(out, cyc, 3); k_bar<<<1, 256>>>(out, cyc, 4); k_bar<<<1, 512>>>(out, cyc, 5);
    k_udiv<<<1, 32>>>(out, cyc, 79); k_bfly<<<1, 32>>>(out, cyc);
    k_round2<<<1, 128>>>(out, cyc, 8); k_round2<<<1, 256>>>(out, cyc, 9); k_round2<<<1, 512>>>(out, cyc, 10);
    k_round1<<<1, 128>>>(out, cyc, 11); k_round1<<<1, 256>>>(out, cyc, 12); k_round1<<<1, 512>>>(out, cyc, 13);
    k_round1<<<1, 32>>>(out, cyc, 14);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    const char *names[] = {"redux.max+xor dep chain", "shfl+max+add dep chain", "ballot+add dep chain", "sts+bar+lds 128thr",
                           "sts+bar+lds 256thr", "sts+bar+lds 512thr", "udiv+mul+add chain", "butterfly argmax 5 steps",
                           "round 2-redux 128thr", "round 2-redux 256thr", "round 2-redux 512thr",
                           "round 1-redux 128thr", "round 1-redux 256thr", "round 1-redux 512thr", "round 1-redux 32thr"};
    for (int i = 0; i < 15; ++i) printf("%-28s %lld cycles/iter\n", names[i], cyc[i]);
    return 0;
}
