// fps.cu -- furthest point sampling, bit-exact with the reference kernel
// (lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173, launch :175-229, init sampling.cpp:70-76).
//
// Design (B200).  The reference runs one 512-thread block per scene that re-reads xyz and
// its `temp` min-distance array from global memory every round and reduces with a 9-level
// __syncthreads tree.  FPS is a chain of m-1 dependent argmax rounds, so what matters is
// the latency of ONE round.  Here a scene is owned by a thread-block cluster (1..16 CTAs):
//   - every point lives in a register slot of exactly one thread for the whole kernel
//     (x, y, z and its running min distance); nothing is re-read from HBM/L2;
//   - a round is: P distance updates per thread -> warp argmax (one redux.sync + one ballot)
//     -> the warp's winner (distance bits + xyz, 16 bytes) goes to shared memory;
//   - single CTA: one bar.sync, every warp reduces the <= 16 warp winners again;
//   - cluster: every warp sends its winner with st.async straight into the shared memory of
//     ALL CTAs of the cluster, completing a transaction mbarrier there: one DSMEM hop per
//     round, no bar.sync and no cluster-wide barrier; every CTA then reduces the same
//     (#CTAs x #warps) table, so all CTAs agree on the next sample without a broadcast.
//   - the output index is not needed by the next round: the owner thread stores its slot
//     number and a parallel pass converts slots to point indices after the last round.
//
// Exactness.  The reference's argmax tie order is an artefact of its strided scan and
// shared-memory tree: among equal maxima the winner is the thread with
// the smallest bit-reversed id, then the smallest k inside that thread.  Points are
// therefore laid out in "rank order" g = bitrev(k mod bs) * cnt + k div bs (bs = the
// reference block size for this n, cnt = ceil(n / bs)), and thread t of the cluster owns
// the contiguous slots [t*P, (t+1)*P).  The reference's winner is then the candidate with
// the largest distance and, among equals, the smallest g -- i.e. the first slot of the
// first lane of the first warp of the first CTA, which is what "strict > in slot order",
// ffs(ballot) and the table order compute.  Distances are non-negative floats, so their
// bit patterns order like unsigned integers; key 0 means "no candidate" (every slot of the
// thread is a skipped point or padding), for which the reference yields index 0.
#include "common.cuh"
#include <stdlib.h>

namespace pn2 {

// fps_bucket.cu: one CTA per scene over spatially bucketed points (large scenes, needs a workspace)
int fps_bucket_min_points(int mode);
size_t fps_bucket_workspace_bytes(int b, int n, int mode);
int fps_bucket_launch(int b, int n, int m, int lg_bs, int cnt, const float *xyz, int pitch, int *idxs, float *new_xyz,
                      float *xyz_copy, void *ws, size_t ws_bytes, long long *prof, cudaStream_t stream);

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
// Optional suspend-time hint of try_wait (-DPN2_FPS_MBAR_HINT=0x989680): the warp would sleep in hardware until the
// phase completes instead of re-issuing the probe (~10 probe + branch pairs per round).  Measured neutral, alone and
// with two CTAs per SM (1.351 / 1.636 ms against 1.353 / 1.630 ms, scripts/gpu_r2_hint.sh): the default probe already
// suspends for about as long as the exchange takes.  Off by default.
#ifndef PN2_FPS_MBAR_HINT
#define PN2_FPS_MBAR_HINT 0
#endif
constexpr uint32_t kMbarSuspendHint = PN2_FPS_MBAR_HINT;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
#if PN2_FPS_MBAR_HINT
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(kMbarSuspendHint)
            : "memory");
#else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
#endif
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
constexpr int kMaxWarps = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kNoSlot = 0xffffffffu;

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

// ---- packed fp32x2 arithmetic (sm_100): two IEEE round-to-nearest operations per instruction, so
// the per-point results are bit-identical to the scalar FADD / FMUL / FFMA sequence of sqdist3.
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Cand {   // a candidate of the argmax: distance and slot offset inside the thread
    float v;
    int i;
};
__device__ __forceinline__ void take_later_if_greater(Cand &a, const Cand &b)
{
    const bool g = b.v > a.v;   // strict: on ties the earlier slot stays
    a.v = g ? b.v : a.v; a.i = g ? b.i : a.i;
}

// P = point slots per thread (even).  REGS: xyz of the slots are kept in registers (P <= 16); otherwise
// they are read back from this CTA's shared-memory copy each round.  CLUSTER: compiled-in switch
// between the single-CTA exchange (shared memory + bar.sync) and the DSMEM exchange.
template <int P, bool REGS, bool CLUSTER, int MAXT, int MINB, int RP = (REGS ? P / 2 : 0)>
__global__ void __launch_bounds__(MAXT, MINB)
fps_kernel(int n, int m, int lg_bs, int cnt, const float *__restrict__ xyz, int pitch, int *__restrict__ idxs,
           float *__restrict__ new_xyz, float *__restrict__ xyz_copy, long long *__restrict__ prof)
{
    static_assert(P % 2 == 0, "slots are processed in pairs");
    extern __shared__ float dyn[];   // sx[P*T], sy[P*T], sz[P*T]: the winner's coordinates are fetched from here
    // winners table, double buffered: (distance key, x, y, z) of every warp of every CTA.  Slot s
    // (= cluster-wide warp number) lives at position (s % K) * 32 + s / K, so that lane l reads its
    // K consecutive slots l*K .. l*K+K-1 at positions i*32 + l: conflict free, and "lowest lane"
    // is "lowest slot".
    __shared__ __align__(16) uint4 table[2][kMaxCluster * kMaxWarps];
    __shared__ __align__(8) uint64_t bar[2];

    // the 40-slot variant is only ever launched with MAXT threads: a compile-time T folds the slot addressing
    const int T = (P == 40) ? MAXT : (int)blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const uint32_t C = CLUSTER ? cluster_nctarank() : 1u, rank = CLUSTER ? cluster_ctarank() : 0u;
    const int scene = blockIdx.x / C;
    const float *p = xyz + (size_t)scene * n * pitch;       // rows of `pitch` floats, xyz first (pitch = 3: plain (n,3))
    float *cp = xyz_copy ? xyz_copy + (size_t)scene * n * 3 : nullptr;   // optional contiguous (n,3) copy, written once
    int *out_idx = idxs + (size_t)scene * m;
    float *out_xyz = new_xyz ? new_xyz + (size_t)scene * m * 3 : nullptr;
    float *sx = dyn, *sy = dyn + P * T, *sz = dyn + 2 * P * T;
    const uint32_t g0 = (rank * T + tid) * P;          // first slot of this thread
    const int entries = (int)C * W;                      // table entries per round
    const int K = (entries + 31) >> 5;                   // entries per lane (1, 2, 4 or 8)
    const int myslot = (int)rank * W + warp;
    const uint32_t mypos = (uint32_t)((myslot % K) * 32 + myslot / K);

    // slot (i, tid) of the shared-memory copy: the two slots of a pair are adjacent (one 8-byte access per pair)
    auto sidx = [&](int i) { return ((i >> 1) * T + tid) * 2 + (i & 1); };
    // RP = slot pairs whose coordinates stay in registers; the others are re-read from shared memory each round
    // (fewer registers per thread -> more scenes resident per SM)
    uint64_t px2[RP > 0 ? RP : 1], py2[RP > 0 ? RP : 1], pz2[RP > 0 ? RP : 1];
    float pt[P];
    const bool vec4 = (pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
#pragma unroll
    for (int i = 0; i < P; i += 2) {
        float c[2][3], t[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = unrank(g0 + i + h, lg_bs, cnt, n);
            c[h][0] = c[h][1] = c[h][2] = 0.f;
            t[h] = -1.f;                                 // -1: never a candidate (fminf keeps it at -1)
            if (k >= 0) {
                if (vec4) {
                    // rows of point_clouds (pitch 132 floats): one 16-byte load per point with a 64-byte L2 fetch hint instead
                    // of three scalar loads -- each used to pull a whole 128-byte line out of DRAM for 12 useful bytes
                    float4 v;
                    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + (size_t)pitch * k));
                    c[h][0] = v.x; c[h][1] = v.y; c[h][2] = v.z;
                } else {
                    c[h][0] = __ldg(p + (size_t)pitch * k);
                    c[h][1] = __ldg(p + (size_t)pitch * k + 1);
                    c[h][2] = __ldg(p + (size_t)pitch * k + 2);
                }
                if (cp) { cp[3 * (size_t)k] = c[h][0]; cp[3 * (size_t)k + 1] = c[h][1]; cp[3 * (size_t)k + 2] = c[h][2]; }
                // sampling_gpu.cu:100-101: float mag compared against the double literal 1e-3
                if (!((double)sqnorm3(c[h][0], c[h][1], c[h][2]) <= 1e-3)) t[h] = 1e10f;   // sampling.cpp:74-76
            }
            sx[sidx(i + h)] = c[h][0]; sy[sidx(i + h)] = c[h][1]; sz[sidx(i + h)] = c[h][2];
        }
        if (i / 2 < RP) {
            px2[i / 2] = pack2(c[0][0], c[1][0]); py2[i / 2] = pack2(c[0][1], c[1][1]); pz2[i / 2] = pack2(c[0][2], c[1][2]);
        }
        pt[i] = t[0]; pt[i + 1] = t[1];
    }
    for (int i = tid; i < 2 * kMaxCluster * kMaxWarps; i += T) (&table[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    // addresses that do not change from round to round
    const uint32_t tab0 = smem_u32(&table[0][0]), bar0 = smem_u32(&bar[0]);
    constexpr uint32_t kTabBytes = kMaxCluster * kMaxWarps * 16;
    uint32_t r_tab = 0, r_bar = 0;                       // lane c < C: this warp's entry / the barrier in CTA c
    if (CLUSTER) {
        if (tid == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (lane < C) {
            r_tab = map_to_cta(tab0 + mypos * 16, lane);
            r_bar = map_to_cta(bar0, lane);
        }
        __syncthreads();
        cluster_sync_all();
    } else {
        __syncthreads();
    }

    // sampling_gpu.cu:85-87: the first sample is point 0, unconditionally
    const float p0x = __ldg(p), p0y = __ldg(p + 1), p0z = __ldg(p + 2);
    float ox = p0x, oy = p0y, oz = p0z;
    if (rank == 0 && tid == 0) {
        out_idx[0] = (int)kNoSlot;                    // converted to index 0 by the final pass
        if (out_xyz) { out_xyz[0] = ox; out_xyz[1] = oy; out_xyz[2] = oz; }
    }

    // optional phase profile (diagnostics): SM cycles of thread 0 of CTA 0 per phase, summed over rounds
    long long pc[5] = {0, 0, 0, 0, 0}, tprev = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (profiling) tprev = clock64();
#define PN2_FPS_MARK(i) if (profiling) { const long long tn = clock64(); pc[i] += tn - tprev; tprev = tn; }

    // The thread's candidate (its largest running distance; earliest slot among equals) is cached across rounds
    // with its coordinates.  Running distances only shrink, and no earlier slot can equal the cached maximum (it
    // would have been the candidate), so the candidate can only change when ITS OWN distance shrinks: one scalar
    // distance per round decides, and the full tournament + coordinate fetch runs only then (every round at the
    // start, a few per cent of the rounds once the samples are dense).
    Cand best;
    float bx, by, bz;
    auto tournament = [&]() {
        Cand cd[P];
#pragma unroll
        for (int i = 0; i < P; ++i) { cd[i].v = pt[i]; cd[i].i = i; }
#pragma unroll
        for (int st = 1; st < P; st *= 2)
#pragma unroll
            for (int i = 0; i + st < P; i += 2 * st) take_later_if_greater(cd[i], cd[i + st]);
        best = cd[0];
        bx = sx[sidx(best.i)]; by = sy[sidx(best.i)]; bz = sz[sidx(best.i)];
    };
    tournament();

    for (int j = 1; j < m; ++j) {
        const int r = j - 1;
        const uint32_t boff = (uint32_t)(r & 1);
        // ---- distance update: all slots independently ----
        const uint64_t nx2 = pack2(-ox, -ox), ny2 = pack2(-oy, -oy), nz2 = pack2(-oz, -oz);
        // does the new sample shrink the candidate's own distance?  (same recipe as the slots: sqdist3)
        // (with few slots per thread the tournament is cheaper than the test: always run it)
        constexpr bool kLazy = P >= 12;
        const bool stale = !kLazy || fminf(sqdist3(bx, by, bz, ox, oy, oz), best.v) != best.v;
#pragma unroll
        for (int i = 0; i < P; i += 2) {
            uint64_t x2, y2, z2;
            if (i / 2 < RP) {
                x2 = px2[i / 2]; y2 = py2[i / 2]; z2 = pz2[i / 2];
            } else {
                x2 = *reinterpret_cast<const uint64_t *>(sx + sidx(i));
                y2 = *reinterpret_cast<const uint64_t *>(sy + sidx(i));
                z2 = *reinterpret_cast<const uint64_t *>(sz + sidx(i));
            }
            // (x - ox)^2 + (y - oy)^2 + (z - oz)^2 as FMUL, FFMA, FFMA with the x term first (sqdist3)
            const uint64_t dx = add2(x2, nx2), dy = add2(y2, ny2), dz = add2(z2, nz2);
            const uint64_t d = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            float d0, d1;
            unpack2(d, d0, d1);
            pt[i] = fminf(d0, pt[i]);
            pt[i + 1] = fminf(d1, pt[i + 1]);
        }
        if (stale) tournament();
        const uint32_t hi = best.v >= 0.f ? __float_as_uint(best.v) + 1u : 0u;
        PN2_FPS_MARK(0)

        // ---- warp winner: largest key, lowest lane among equals ----
        const uint32_t wmax = __reduce_max_sync(kFull, hi);
        const int wsrc = __ffs(__ballot_sync(kFull, hi == wmax)) - 1;
        const bool warp_owner = lane == wsrc;
        PN2_FPS_MARK(1)
        const uint32_t tab = tab0 + boff * kTabBytes;
        if (!CLUSTER) {
            if (warp_owner)
                sts128(tab + mypos * 16, make_uint4(wmax, __float_as_uint(bx), __float_as_uint(by), __float_as_uint(bz)));
            __syncthreads();
        } else {
            // every warp publishes its winner to all CTAs of the cluster (lane c -> CTA c)
            const uint32_t vx = __shfl_sync(kFull, __float_as_uint(bx), wsrc);
            const uint32_t vy = __shfl_sync(kFull, __float_as_uint(by), wsrc);
            const uint32_t vz = __shfl_sync(kFull, __float_as_uint(bz), wsrc);
            if (lane < C) st_async_v4(r_tab + boff * kTabBytes, wmax, vx, vy, vz, r_bar + boff * 8);
            if (tid == 0) mbar_arrive_expect_tx(bar0 + boff * 8, 16u * entries);
            mbar_wait(bar0 + boff * 8, (r >> 1) & 1);
        }
        PN2_FPS_MARK(2)
        // ---- winner of the whole scene: every warp reduces the same table ----
        // lane l holds table positions i*32 + l, i < K: keep the largest key, the earliest position among equals
        uint4 e0 = lds128(tab + (uint32_t)lane * 16);
        int e0i = 0;
        for (int i = 1; i < K; ++i) {
            const uint4 ei = lds128(tab + (uint32_t)(i * 32 + lane) * 16);
            if (ei.x > e0.x) { e0 = ei; e0i = i; }
        }
        const uint32_t smax = __reduce_max_sync(kFull, e0.x);
        const int ssrc = __ffs(__ballot_sync(kFull, e0.x == smax)) - 1;   // lowest lane = lowest slot
        const float nx = __uint_as_float(__shfl_sync(kFull, e0.y, ssrc));
        const float ny = __uint_as_float(__shfl_sync(kFull, e0.z, ssrc));
        const float nz = __uint_as_float(__shfl_sync(kFull, e0.w, ssrc));
        const int sslot = ssrc * K + __shfl_sync(kFull, e0i, ssrc);

        const bool none = smax == 0u;   // every point skipped: the reference's reduction leaves besti = 0
        ox = none ? p0x : nx; oy = none ? p0y : ny; oz = none ? p0z : nz;
        PN2_FPS_MARK(3)
        // the thread that owns the winning slot records it (off the critical path)
        if (warp_owner && sslot == myslot) {
            out_idx[j] = none ? (int)kNoSlot : (int)(g0 + best.i);
            if (out_xyz) { out_xyz[3 * j] = ox; out_xyz[3 * j + 1] = oy; out_xyz[3 * j + 2] = oz; }
        }
    }
    if (profiling) {
        PN2_FPS_MARK(4)
        for (int i = 0; i < 5; ++i) prof[i] = pc[i];
    }
#undef PN2_FPS_MARK
    // slots -> point indices (the stores above were made by threads of this cluster)
    __threadfence();
    if (CLUSTER) cluster_sync_all();   // also: no CTA leaves while a peer may still address its shared memory
    else __syncthreads();
    for (int j = rank * T + tid; j < m; j += C * T) {
        const uint32_t g = (uint32_t)out_idx[j];
        out_idx[j] = g == kNoSlot ? 0 : unrank(g, lg_bs, cnt, n);
    }
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

static const int kPpt[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40};

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
        // 20 slots per thread: a 40k-point scene takes 8 CTAs.  Measured on B200 the round latency of 8 and 16
        // CTAs is the same within noise (0.85-1.0 us), and the smaller cluster leaves room for twice as many
        // scenes in flight (throughput with several batches on different streams: +15 %).
        threads = 256;
        cluster = 2;
        while (cluster < kMaxCluster && slots > (long long)cluster * threads * 20) cluster *= 2;
        // Four warps per CTA with 40 slots per thread (two CTAs per SM at 246 registers) when the scene fills such a
        // cluster: half the warps per SM and a 32-entry winners table (one entry per lane).  Measured at 40 000
        // points, cluster 8: 0.666 us per round alone and 0.388 ms per 8-scene batch with eight batches in flight,
        // against 0.675 us and 0.438 ms for 256 threads x 20 slots.
        for (int c = 2; c <= kMaxCluster; c *= 2)
            if (slots <= (long long)c * 128 * 40 && slots > (long long)c * 128 * 32) { cluster = c; threads = 128; break; }
        if (slots > (long long)cluster * threads * 40) threads = 512;
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
    long long per = (slots + (long long)cluster * threads - 1) / ((long long)cluster * threads);
    // 40 slots per thread exist only for 128-thread clusters (dispatch): wider CTAs take at most 32
    if (per > 32 && threads == 256) { threads = 512; per = (slots + (long long)cluster * threads - 1) / ((long long)cluster * threads); }
    if (per > 32 && !(threads == 128 && cluster > 1)) return false;
    int ppt = -1;
    for (int v : kPpt)
        if (v >= per) { ppt = v; break; }
    if (ppt < 0) return false;
    pl->cluster = cluster; pl->threads = threads; pl->ppt = ppt;
    return true;
}

template <int P, bool REGS, int MAXT>
static int launch(const FpsPlan &pl, int b, int n, int m, const float *xyz, int pitch, int *idxs, float *new_xyz,
                  float *xyz_copy, long long *prof, cudaStream_t stream)
{
    // large clusters of 256-thread CTAs: a second variant capped at 128 registers lets two CTAs share an SM
    // (twice the scenes in flight when several batches run concurrently)
    static const bool two_per_sm = [] { const char *e = getenv("PN2_FPS_MINB"); return !e || atoi(e) == 2; }();
    auto kern = pl.cluster > 1 ? ((MAXT <= 256 && P >= 16 && two_per_sm) ? fps_kernel<P, REGS, true, MAXT, (MAXT <= 256 && P >= 16) ? 2 : 1>
                                                                         : fps_kernel<P, REGS, true, MAXT, 1>)
                               : fps_kernel<P, REGS, false, MAXT, 1>;
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
    PN2_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, n, m, pl.lg_bs, pl.cnt, xyz, pitch, idxs, new_xyz, xyz_copy, prof));
    count_launches(1);
    return PN2_OK;
}

static int dispatch(const FpsPlan &pl, int b, int n, int m, const float *xyz, int pitch, int *idxs,
                    float *new_xyz, float *xyz_copy, long long *prof, cudaStream_t s)
{
#define PN2_FPS_CASE(P, REGS, MAXT) \
    case P: return launch<P, REGS, MAXT>(pl, b, n, m, xyz, pitch, idxs, new_xyz, xyz_copy, prof, s)
    if (pl.threads == 128 && pl.ppt == 40 && pl.cluster > 1) {
        // PN2_FPS_RP: slot pairs whose coordinates stay in registers (default 6 of 20; the others are re-read from the
        // CTA's shared-memory copy every round).  164 registers instead of 237 -> three CTAs per SM, 48 scenes resident
        // instead of 32.  Measured (scripts/fps_sat_one.py, gpu_r2_check6.sh): alone 1.38 against 1.35 ms, saturated
        // 0.356 against 0.409 ms per 8 scenes, bench step +5-8 %.  20 = all in registers (two CTAs per SM).
        // A single scene is a latency problem (nothing else competes for its 8 SMs): all-register kernel.
        static const int rp_env = [] { const char *e = getenv("PN2_FPS_RP"); return e ? atoi(e) : -1; }();
        const int rp = rp_env >= 0 ? rp_env : (b == 1 ? 20 : 6);
        auto kern = rp == 0 ? fps_kernel<40, false, true, 128, 3, 0>
                  : rp == 4 ? fps_kernel<40, false, true, 128, 3, 4>
                  : rp == 6 ? fps_kernel<40, false, true, 128, 3, 6>
                  : rp == 8 ? fps_kernel<40, false, true, 128, 3, 8>
                            : fps_kernel<40, true, true, 128, 2>;
        const size_t smem = (size_t)3 * 40 * 128 * sizeof(float);
        PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)b * pl.cluster); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pl.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PN2_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, n, m, pl.lg_bs, pl.cnt, xyz, pitch, idxs, new_xyz, xyz_copy, prof));
        count_launches(1);
        return PN2_OK;
    }
    if (pl.threads <= 256) {
        switch (pl.ppt) {
            PN2_FPS_CASE(2, true, 256); PN2_FPS_CASE(4, true, 256); PN2_FPS_CASE(6, true, 256);
            PN2_FPS_CASE(8, true, 256); PN2_FPS_CASE(10, true, 256); PN2_FPS_CASE(12, true, 256);
            PN2_FPS_CASE(16, true, 256); PN2_FPS_CASE(20, true, 256); PN2_FPS_CASE(24, true, 256);
            PN2_FPS_CASE(32, true, 256);
        }
    } else {
        switch (pl.ppt) {
            PN2_FPS_CASE(2, true, 512); PN2_FPS_CASE(4, true, 512); PN2_FPS_CASE(6, true, 512);
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

extern "C" size_t pn2_furthest_point_sampling_workspace_bytes(int b, int n, int)
{
    return b > 0 ? fps_bucket_workspace_bytes(b, n, PN2_FPS_LATENCY) : 0;
}

extern "C" size_t pn2_furthest_point_sampling_workspace_bytes_mode(int b, int n, int, int mode)
{
    if (mode != PN2_FPS_LATENCY && mode != PN2_FPS_THROUGHPUT) return 0;
    return b > 0 ? fps_bucket_workspace_bytes(b, n, mode) : 0;
}

static int fps_entry(int b, int n, int m, const float *xyz, int pitch, int *idxs, float *new_xyz, float *xyz_copy,
                     long long *prof, pn2_stream_t stream, void *ws = nullptr, size_t ws_bytes = 0)
{
    if (b < 0 || n < 0 || m < 0 || pitch < 3) return PN2_ERR_INVALID_ARGUMENT;
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
    // large scenes with a workspace: spatially bucketed points, one CTA per scene (fps_bucket.cu)
    // The caller chooses by handing over a workspace: the size query of its mode returned non-zero
    const size_t need = fps_bucket_workspace_bytes(b, n, PN2_FPS_THROUGHPUT);
    if (ws && need && ws_bytes >= need) {
        const int lg = ref_block_lg(n), bs = 1 << lg;
        return fps_bucket_launch(b, n, m, lg, (n + bs - 1) / bs, xyz, pitch, idxs, new_xyz, xyz_copy, ws, ws_bytes, prof,
                                 as_stream(stream));
    }
    FpsPlan pl;
    if (!make_plan(n, &pl)) return PN2_ERR_INVALID_ARGUMENT;   // n > 16 CTAs x 512 threads x 40 slots: needs the workspace
    return dispatch(pl, b, n, m, xyz, pitch, idxs, new_xyz, xyz_copy, prof, as_stream(stream));
}

extern "C" int pn2_furthest_point_sampling_xyz(int b, int n, int m, const float *xyz, int *idxs,
                                               float *new_xyz, pn2_stream_t stream)
{
    return fps_entry(b, n, m, xyz, 3, idxs, new_xyz, nullptr, nullptr, stream);
}

// Same over rows of `pitch` floats whose first three are x, y, z (point_clouds read in place); xyz_copy (b,n,3), when
// given, receives the contiguous coordinates the later kernels of the step read (written once while the points load).
extern "C" int pn2_furthest_point_sampling_rows(int b, int n, int m, const float *rows, int pitch, int *idxs,
                                                float *new_xyz, float *xyz_copy, pn2_stream_t stream)
{
    return fps_entry(b, n, m, rows, pitch, idxs, new_xyz, xyz_copy, nullptr, stream);
}

// Diagnostic: same launch, and prof[0..4] (device, 5 x int64) receives the SM cycles thread 0 of CTA 0
// spent in {distance update, warp argmax, exchange (barrier / DSMEM wait), scene argmax, -} summed over rounds.
extern "C" int pn2_debug_fps_profile(int b, int n, int m, const float *xyz, int *idxs, long long *prof,
                                     pn2_stream_t stream)
{
    if (!prof) return PN2_ERR_INVALID_ARGUMENT;
    return fps_entry(b, n, m, xyz, 3, idxs, nullptr, nullptr, prof, stream);
}

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float *xyz, void *workspace, size_t workspace_bytes,
                                           int *idxs, pn2_stream_t stream)
{
    return fps_entry(b, n, m, xyz, 3, idxs, nullptr, nullptr, nullptr, stream, workspace, workspace_bytes);
}

extern "C" int pn2_furthest_point_sampling_xyz_ws(int b, int n, int m, const float *xyz, int *idxs, float *new_xyz,
                                                  void *workspace, size_t workspace_bytes, pn2_stream_t stream)
{
    return fps_entry(b, n, m, xyz, 3, idxs, new_xyz, nullptr, nullptr, stream, workspace, workspace_bytes);
}

extern "C" int pn2_furthest_point_sampling_rows_ws(int b, int n, int m, const float *rows, int pitch, int *idxs,
                                                   float *new_xyz, float *xyz_copy, void *workspace, size_t workspace_bytes,
                                                   pn2_stream_t stream)
{
    return fps_entry(b, n, m, rows, pitch, idxs, new_xyz, xyz_copy, nullptr, stream, workspace, workspace_bytes);
}

// Diagnostic for the bucketed kernel: prof (device, 6 x int64) = SM cycles of thread 0 of CTA 0 in {bucket tests,
// bucket updates, thread/warp argmax, barrier, table argmax, -} summed over the rounds.
extern "C" int pn2_debug_fps_bucket_profile(int b, int n, int m, const float *xyz, int *idxs, long long *prof,
                                            void *workspace, size_t workspace_bytes, pn2_stream_t stream)
{
    if (!prof) return PN2_ERR_INVALID_ARGUMENT;
    return fps_entry(b, n, m, xyz, 3, idxs, nullptr, nullptr, prof, stream, workspace, workspace_bytes);
}
