// fps.cu -- furthest point sampling, bit-exact with the reference kernel
// (lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173, launch :175-229, init sampling.cpp:70-76).
//
// Design (B200).  The reference runs one 512-thread block per scene that re-reads xyz and
// its `temp` min-distance array from global memory every round and reduces with a 9-level
// __syncthreads tree.  FPS is a chain of m-1 dependent argmax rounds, so what matters is
// the latency of ONE round.  Here a scene is owned by a thread-block cluster (1..16 CTAs):
//   - every point lives in a register slot of exactly one thread for the whole kernel
//     (x, y, z and its running min distance), nothing is re-read from HBM/L2;
//   - a round is: P distance updates per thread -> warp argmax with two redux.sync ->
//     per-warp winners in shared memory, one bar.sync -> CTA winner;
//   - CTAs exchange their winner (key + xyz, 20 bytes) with st.async straight into every
//     peer's shared memory, completing a transaction mbarrier there: one DSMEM hop per
//     round, no cluster-wide barrier.
//
// Exactness.  The reference's argmax tie order is an artefact of its strided scan and
// shared-memory tree: among equal maxima the winner is the thread with
// the smallest bit-reversed id, then the smallest k inside that thread.  Points are
// therefore laid out in "rank order" g = bitrev(k mod bs) * cnt + k div bs (bs = the
// reference block size for this n, cnt = ceil(n / bs)), so that the reference's winner is
// simply the candidate with the largest distance and, among equals, the smallest g.
// Distances are non-negative floats, so their bit patterns order like unsigned integers
// and the argmax is an integer max over (dist_bits + 1, ~g); key 0 means "no candidate"
// (every point of the thread is skipped or a padding slot), for which the reference
// yields index 0.
#include "common.cuh"
#include <stdlib.h>

namespace pn2 {

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 16-byte / 4-byte store into a peer CTA's shared memory that also completes `bytes` on the
// peer's transaction barrier.
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t rbar)
{
    asm volatile(
        "st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
        ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
        : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(a), "r"(rbar)
                 : "memory");
}

constexpr int kMaxCluster = 16;
constexpr unsigned kFull = 0xffffffffu;

// rank-order slot g -> original point index k (or -1 for a padding slot)
__device__ __forceinline__ int unrank(uint32_t g, int lg_bs, int cnt, int n)
{
    const uint32_t bt = g / (uint32_t)cnt;
    const uint32_t q = g - bt * (uint32_t)cnt;
    if (bt >= (1u << lg_bs)) return -1;
    const uint32_t t = lg_bs == 0 ? 0u : (__brev(bt) >> (32 - lg_bs));
    const long long k = (long long)q * (1ll << lg_bs) + t;
    return k < n ? (int)k : -1;
}

// P = point slots per thread.  REGS: xyz of the slots are kept in registers (P <= 16);
// otherwise they are read back from this CTA's shared-memory copy each round.
template <int P, bool REGS, int MAXT>
__global__ void __launch_bounds__(MAXT)
fps_kernel(int n, int m, int lg_bs, int cnt, const float *__restrict__ xyz, int *__restrict__ idxs,
           float *__restrict__ new_xyz)
{
    extern __shared__ float dyn[];   // sx[P*T], sy[P*T], sz[P*T]
    __shared__ uint2 w_key[2][32];
    __shared__ float4 w_xyz[2][32];
    __shared__ __align__(16) uint4 c_msg[2][kMaxCluster];   // (hi, lo, x, y) of each CTA's winner
    __shared__ uint32_t c_z[2][kMaxCluster];
    __shared__ __align__(8) uint64_t c_bar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
    const int scene = blockIdx.x / C;
    const float *p = xyz + (size_t)scene * n * 3;
    int *out_idx = idxs + (size_t)scene * m;
    float *out_xyz = new_xyz ? new_xyz + (size_t)scene * m * 3 : nullptr;
    float *sx = dyn, *sy = dyn + P * T, *sz = dyn + 2 * P * T;
    const uint32_t gstride = C * T, gbase = rank * T + tid;

    float px[P], py[P], pz[P], pt[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = unrank(i * gstride + gbase, lg_bs, cnt, n);
        float x = 0.f, y = 0.f, z = 0.f, t = -1.f;   // -1: never a candidate (fminf keeps it at -1)
        if (k >= 0) {
            x = __ldg(p + 3 * (size_t)k);
            y = __ldg(p + 3 * (size_t)k + 1);
            z = __ldg(p + 3 * (size_t)k + 2);
            // sampling_gpu.cu:100-101: float mag compared against the double literal 1e-3
            if (!((double)sqnorm3(x, y, z) <= 1e-3)) t = 1e10f;   // sampling.cpp:74-76
        }
        px[i] = x; py[i] = y; pz[i] = z; pt[i] = t;
        sx[i * T + tid] = x; sy[i * T + tid] = y; sz[i * T + tid] = z;
    }
    if (C > 1) {
        if (tid == 0) {
            mbar_init(smem_u32(&c_bar[0]), 1);
            mbar_init(smem_u32(&c_bar[1]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();
    }

    // sampling_gpu.cu:85-87: the first sample is point 0, unconditionally
    float ox = __ldg(p), oy = __ldg(p + 1), oz = __ldg(p + 2);
    if (rank == 0 && tid == 0) {
        out_idx[0] = 0;
        if (out_xyz) { out_xyz[0] = ox; out_xyz[1] = oy; out_xyz[2] = oz; }
    }

    for (int j = 1; j < m; ++j) {
        const int r = j - 1, buf = r & 1;
        float best = -1.f;
        int bi = 0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            float x, y, z;
            if (REGS) { x = px[i]; y = py[i]; z = pz[i]; }
            else { x = sx[i * T + tid]; y = sy[i * T + tid]; z = sz[i * T + tid]; }
            const float d = sqdist3(x, y, z, ox, oy, oz);
            const float d2 = fminf(d, pt[i]);
            pt[i] = d2;
            if (d2 > best) { best = d2; bi = i; }   // strict: earliest slot wins inside a thread
        }
        const uint32_t hi = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        const uint32_t lo = ~(uint32_t)(bi * gstride + gbase);

        // warp winner
        uint32_t mh = __reduce_max_sync(kFull, hi);
        uint32_t ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
        if (hi == mh && lo == ml) {
            const int s = bi * T + tid;
            w_key[buf][warp] = make_uint2(mh, ml);
            w_xyz[buf][warp] = make_float4(sx[s], sy[s], sz[s], 0.f);
        }
        __syncthreads();
        // CTA winner (every warp computes it redundantly: no second barrier)
        const uint2 wk = lane < W ? w_key[buf][lane] : make_uint2(0u, 0u);
        mh = __reduce_max_sync(kFull, wk.x);
        ml = __reduce_max_sync(kFull, wk.x == mh ? wk.y : 0u);
        int src = __ffs(__ballot_sync(kFull, wk.x == mh && wk.y == ml)) - 1;
        const float4 wv = w_xyz[buf][src];
        float nx = wv.x, ny = wv.y, nz = wv.z;

        if (C > 1) {
            // cluster winner: all-to-all of 20-byte messages through DSMEM
            const uint32_t bar = smem_u32(&c_bar[buf]);
            if (warp == 0 && lane < C) {
                const uint32_t rbar = map_to_cta(bar, lane);
                st_async_v4(map_to_cta(smem_u32(&c_msg[buf][rank]), lane), mh, ml,
                            __float_as_uint(nx), __float_as_uint(ny), rbar);
                st_async_b32(map_to_cta(smem_u32(&c_z[buf][rank]), lane), __float_as_uint(nz), rbar);
            }
            if (tid == 0) mbar_arrive_expect_tx(bar, 20u * C);
            mbar_wait(bar, (r >> 1) & 1);
            const uint4 e = lane < C ? c_msg[buf][lane] : make_uint4(0u, 0u, 0u, 0u);
            mh = __reduce_max_sync(kFull, e.x);
            ml = __reduce_max_sync(kFull, e.x == mh ? e.y : 0u);
            src = __ffs(__ballot_sync(kFull, e.x == mh && e.y == ml)) - 1;
            nx = __uint_as_float(__shfl_sync(kFull, e.z, src));
            ny = __uint_as_float(__shfl_sync(kFull, e.w, src));
            nz = __uint_as_float(c_z[buf][src]);
        }

        int k = 0;
        if (mh != 0u) {
            k = unrank(~ml, lg_bs, cnt, n);
            ox = nx; oy = ny; oz = nz;
        } else {
            // every point skipped: the reference's reduction leaves besti = 0 (sampling_gpu.cu:93-94,170)
            ox = __ldg(p); oy = __ldg(p + 1); oz = __ldg(p + 2);
        }
        if (rank == 0 && tid == 0) {
            out_idx[j] = k;
            if (out_xyz) { out_xyz[3 * j] = ox; out_xyz[3 * j + 1] = oy; out_xyz[3 * j + 2] = oz; }
        }
    }
    if (C > 1) cluster_sync_all();   // no CTA leaves while a peer may still address its shared memory
}

struct FpsPlan {
    int cluster, threads, ppt, lg_bs, cnt;
};

// cuda_utils.h:13-19 opt_n_threads(): 2^floor(log2 n) clamped to [1, 512]
static int ref_block_lg(int n)
{
    int lg = 0;
    while ((2 << lg) <= n && lg < 9) ++lg;
    return lg;
}

static const int kPpt[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};

static bool make_plan(int n, FpsPlan *pl)
{
    pl->lg_bs = ref_block_lg(n);
    const int bs = 1 << pl->lg_bs;
    pl->cnt = (n + bs - 1) / bs;
    const long long slots = (long long)bs * pl->cnt;
    int cluster, threads;
    if (slots <= 1024) { cluster = 1; threads = 128; }
    else if (slots <= 4096) { cluster = 1; threads = 256; }
    else if (slots <= 8192) { cluster = 1; threads = 512; }
    else {
        threads = 512;
        cluster = 2;
        while (cluster < kMaxCluster && slots > (long long)cluster * threads * 10) cluster *= 2;
    }
    // tuning overrides (benchmark sweeps): PN2_FPS_CLUSTER in {1,2,4,8,16}, PN2_FPS_THREADS in {128,256,512}
    if (const char *e = getenv("PN2_FPS_CLUSTER")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) cluster = v;
    }
    if (const char *e = getenv("PN2_FPS_THREADS")) {
        const int v = atoi(e);
        if (v == 128 || v == 256 || v == 512) threads = v;
    }
    const long long per = (slots + (long long)cluster * threads - 1) / ((long long)cluster * threads);
    int ppt = -1;
    for (int v : kPpt)
        if (v >= per) { ppt = v; break; }
    if (ppt < 0) return false;
    pl->cluster = cluster; pl->threads = threads; pl->ppt = ppt;
    return true;
}

template <int P, bool REGS, int MAXT>
static int launch(const FpsPlan &pl, int b, int n, int m, const float *xyz, int *idxs, float *new_xyz,
                  cudaStream_t stream)
{
    auto kern = fps_kernel<P, REGS, MAXT>;
    const size_t smem = (size_t)3 * P * pl.threads * sizeof(float);
    PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (pl.cluster > 8)
        PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * pl.cluster);
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PN2_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, n, m, pl.lg_bs, pl.cnt, xyz, idxs, new_xyz));
    return PN2_OK;
}

static int dispatch(const FpsPlan &pl, int b, int n, int m, const float *xyz, int *idxs,
                    float *new_xyz, cudaStream_t s)
{
#define PN2_FPS_CASE(P, REGS, MAXT) \
    case P: return launch<P, REGS, MAXT>(pl, b, n, m, xyz, idxs, new_xyz, s)
    if (pl.threads <= 256) {
        switch (pl.ppt) {
            PN2_FPS_CASE(1, true, 256); PN2_FPS_CASE(2, true, 256); PN2_FPS_CASE(3, true, 256);
            PN2_FPS_CASE(4, true, 256); PN2_FPS_CASE(5, true, 256); PN2_FPS_CASE(6, true, 256);
            PN2_FPS_CASE(8, true, 256); PN2_FPS_CASE(10, true, 256); PN2_FPS_CASE(12, true, 256);
            PN2_FPS_CASE(16, true, 256); PN2_FPS_CASE(20, true, 256); PN2_FPS_CASE(24, true, 256);
            PN2_FPS_CASE(32, true, 256);
        }
    } else {
        switch (pl.ppt) {
            PN2_FPS_CASE(1, true, 512); PN2_FPS_CASE(2, true, 512); PN2_FPS_CASE(3, true, 512);
            PN2_FPS_CASE(4, true, 512); PN2_FPS_CASE(5, true, 512); PN2_FPS_CASE(6, true, 512);
            PN2_FPS_CASE(8, true, 512); PN2_FPS_CASE(10, true, 512); PN2_FPS_CASE(12, true, 512);
            PN2_FPS_CASE(16, true, 512); PN2_FPS_CASE(20, false, 512); PN2_FPS_CASE(24, false, 512);
            PN2_FPS_CASE(32, false, 512);
        }
    }
#undef PN2_FPS_CASE
    return PN2_ERR_INVALID_ARGUMENT;
}

}  // namespace pn2

using namespace pn2;

extern "C" size_t pn2_furthest_point_sampling_workspace_bytes(int, int, int) { return 0; }

extern "C" int pn2_furthest_point_sampling_xyz(int b, int n, int m, const float *xyz, int *idxs,
                                               float *new_xyz, pn2_stream_t stream)
{
    if (b < 0 || n < 0 || m < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || m == 0) return PN2_OK;
    if (!idxs) return PN2_ERR_INVALID_ARGUMENT;
    if (n == 0) {
        // the reference launches over an empty scene and leaves its zero-filled output untouched
        PN2_CUDA_TRY(cudaMemsetAsync(idxs, 0, sizeof(int) * (size_t)b * m, as_stream(stream)));
        if (new_xyz)
            PN2_CUDA_TRY(cudaMemsetAsync(new_xyz, 0, sizeof(float) * 3 * (size_t)b * m, as_stream(stream)));
        return PN2_OK;
    }
    if (!xyz) return PN2_ERR_INVALID_ARGUMENT;
    FpsPlan pl;
    if (!make_plan(n, &pl)) return PN2_ERR_INVALID_ARGUMENT;   // n > 16 CTAs x 512 threads x 32 slots
    return dispatch(pl, b, n, m, xyz, idxs, new_xyz, as_stream(stream));
}

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float *xyz, void *, size_t,
                                           int *idxs, pn2_stream_t stream)
{
    return pn2_furthest_point_sampling_xyz(b, n, m, xyz, idxs, nullptr, stream);
}
