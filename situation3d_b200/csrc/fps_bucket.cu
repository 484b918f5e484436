// fps_bucket.cu -- furthest point sampling for large scenes: ONE CTA per scene, points binned into spatial
// buckets, exact pruning of the distance update.  Bit-exact with the reference kernel
// (lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173, init sampling.cpp:70-76).
//
// Why.  FPS is m-1 dependent rounds: temp[k] = min(temp[k], |p_k - last|^2) for every point, then an argmax.
// The cluster kernel of fps.cu keeps every point in a register and pays a DSMEM exchange per round over 8 CTAs
// (~1300 cycles per round, four SMs' worth of registers per scene).  But a round only CHANGES the points closer
// to the new sample than their running distance -- after a few dozen samples that is a small neighbourhood.
//
// How.  The prologue bins the scene's points into a 32^3 Morton grid (shared-memory histogram, counting sort
// into a workspace; the order inside a cell is irrelevant to the result) and cuts the sorted list into buckets of
// 32*PPL points; CH consecutive buckets form a super-bucket owned by one thread, which keeps in REGISTERS the
// bounding boxes, each bucket's largest running distance and that candidate's index and coordinates.  A round:
//   1. every thread tests its super-bucket (then its CH buckets) against the new sample: with
//      lb = |max(lo - s, s - hi, 0)|^2 evaluated by the same FSUB/FMUL/FFMA/FFMA recipe as the point distances,
//      lb <= d_k holds in fp32 for every point k of the box (rounding is monotonic), so lb >= bucket max means
//      min(d_k, temp[k]) == temp[k] for all of them: the bucket is skipped EXACTLY (NaN / inf coordinates never
//      lower a distance in the reference either: fminf drops a NaN, and they are left out of the boxes);
//   2. the owning warp streams each surviving bucket (float4 x,y,z,index per point from L2, running distances in
//      shared memory), updates the distances and, if any changed, re-derives the bucket's candidate
//      (redux.sync max on the distance bits; ties by the reference's order, below);
//   3. argmax over the threads' cached candidates: warp redux -> 16-entry table -> one bar.sync -> every warp
//      reduces the table.
// About 35 full passes' worth of point updates for 2047 rounds over 40 000 points instead of 2047 passes.
//
// Tie order.  The reference's strided scan + shared-memory tree picks, among equal maxima, the smallest
// "rank" g(k) = bitrev(k mod bs) * cnt + k div bs (fps.cu).  Equal keys are rare (duplicate
// points), so every comparison is on the distance bits first and evaluates g only on a tie.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace pn2 {
namespace {

constexpr int kBT = 512;              // threads per CTA
constexpr int kBW = kBT / 32;         // warps
constexpr int kCellBits = 5;
constexpr int kCells = 1 << (3 * kCellBits);
constexpr int kCellsPerWarp = kCells / kBW;
constexpr unsigned kFullMask = 0xffffffffu;
constexpr uint32_t kSkipCell = 0xffffu;
constexpr size_t kMaxSmem = 227 * 1024 - 2048;

// order-preserving map float -> uint32 (for redux.sync min / max over signed floats)
__device__ __forceinline__ uint32_t ord_key(float f)
{
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord_inv(uint32_t k)
{
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}
__device__ __forceinline__ uint32_t spread5(uint32_t v)   // bit i -> bit 3i (5 bits)
{
    v = (v | (v << 8)) & 0x100fu;
    v = (v | (v << 4)) & 0x10c3u;
    v = (v | (v << 2)) & 0x1249u;
    return v;
}
__device__ __forceinline__ uint32_t rank_of(uint32_t k, int lg_bs, int cnt)
{
    const uint32_t t = k & ((1u << lg_bs) - 1u);
    const uint32_t bt = lg_bs ? (__brev(t) >> (32 - lg_bs)) : 0u;
    return bt * (uint32_t)cnt + (k >> lg_bs);
}
__device__ __forceinline__ bool finite3(float x, float y, float z)
{
    return fabsf(x) < INFINITY && fabsf(y) < INFINITY && fabsf(z) < INFINITY;
}
// squared distance from s to the box [lo, hi], a lower bound (in fp32, same recipe as sqdist3) of the
// squared distance from s to any point inside the box
__device__ __forceinline__ float box_lb(float lx, float ly, float lz, float hx, float hy, float hz, float sx, float sy, float sz)
{
    const float dx = fmaxf(fmaxf(__fsub_rn(lx, sx), __fsub_rn(sx, hx)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(ly, sy), __fsub_rn(sy, hy)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lz, sz), __fsub_rn(sz, hz)), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
__device__ __forceinline__ uint32_t dist_key(float d) { return d >= 0.f ? __float_as_uint(d) + 1u : 0u; }
__device__ __forceinline__ float key_dist(uint32_t k) { return k ? __uint_as_float(k - 1u) : -1.f; }

struct Cand {            // argmax candidate: key (0 = none), original point index, coordinates
    uint32_t key, idx;
    float x, y, z;
};
// a > b in the reference's order (larger distance, then smaller rank); evaluates ranks only on a tie
__device__ __forceinline__ bool cand_better(uint32_t ak, uint32_t ai, uint32_t bk, uint32_t bi, int lg_bs, int cnt)
{
    if (ak != bk) return ak > bk;
    if (ak == 0u) return false;
    return rank_of(ai, lg_bs, cnt) < rank_of(bi, lg_bs, cnt);
}
// warp argmax of per-lane candidates; the result is in every lane
__device__ __forceinline__ Cand warp_best(const Cand &c, int lg_bs, int cnt)
{
    const uint32_t kmax = __reduce_max_sync(kFullMask, c.key);
    const unsigned eqm = __ballot_sync(kFullMask, c.key == kmax);
    int src = __ffs(eqm) - 1;
    if (kmax != 0u && (eqm & (eqm - 1u))) {      // several lanes hold the maximum: smallest rank wins
        const uint32_t r = c.key == kmax ? rank_of(c.idx, lg_bs, cnt) : 0xffffffffu;
        const uint32_t rm = __reduce_min_sync(kFullMask, r);
        src = __ffs(__ballot_sync(kFullMask, r == rm)) - 1;
    }
    Cand w;
    w.key = kmax;
    w.idx = __shfl_sync(kFullMask, c.idx, src);
    w.x = __shfl_sync(kFullMask, c.x, src);
    w.y = __shfl_sync(kFullMask, c.y, src);
    w.z = __shfl_sync(kFullMask, c.z, src);
    return w;
}

// PPL points per lane in a bucket (bucket = 32*PPL points), CH buckets per super-bucket, SS super-buckets per
// thread; SDIST: running distances in shared memory (else in the workspace).  The CTA starts with kBT threads for the
// binning; the rounds run on the first `nwa` warps only (one thread per super-bucket), the others leave.
// PHASE 0: the whole kernel.  PHASE 1 / 2: the same code as two launches -- the binning (kBT threads, 128 KB histogram)
// and the rounds (only the `nwa` warps that run them, no histogram: <= kSplitThreads threads at <= 102 registers and
// ~30 KB of shared memory, so that TWO scenes share an SM; the rounds are barrier- and latency-bound, issue slots a
// third busy).  The count of competing points travels through the workspace.
// PHASE 3: one launch again, but of kSplitThreads threads from the start and with the histogram in the workspace (L2
// atomics instead of shared memory): the binning is slower, the whole kernel fits two scenes per SM.
constexpr int kSplitThreads = 320;
template <int PPL, int CH, int SS, bool SDIST, int PHASE = 0>
__global__ void __launch_bounds__(PHASE >= 2 ? kSplitThreads : kBT, PHASE >= 2 ? 2 : 1)
fps_bucket_kernel(int n, int m, int lg_bs, int cnt, const float *__restrict__ xyz, int pitch, int *__restrict__ idxs,
                  float *__restrict__ new_xyz, float *__restrict__ xyz_copy, unsigned char *__restrict__ ws,
                  size_t ws_stride, int npad, int nwa, long long *__restrict__ prof)
{
    static_assert(CH == 4, "the work-queue bitmaps hold four buckets per lane");
    constexpr int BS = 32 * PPL;
    constexpr int kWords = SS * CH;                          // bitmap words per warp: one per (ss, c), bit = owner lane
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(dyn);     // prologue: cell histogram / offsets (PHASE 3: in the workspace)
    __shared__ uint32_t red[kBW][8];
    __shared__ uint32_t wbase[kBW + 1];
    __shared__ __align__(8) int2 tableA[kBW];                // per warp: best (distance bits, bucket) it accounts for this round
    __shared__ __align__(16) uint32_t qmask[2][kBW * kWords];   // the round's work queue: which buckets can change (by round parity)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int scene = blockIdx.x;
    const float *p = xyz + (size_t)scene * n * pitch;
    float *cp = xyz_copy ? xyz_copy + (size_t)scene * n * 3 : nullptr;
    int *out_idx = idxs + (size_t)scene * m;
    float *out_xyz = new_xyz ? new_xyz + (size_t)scene * m * 3 : nullptr;
    unsigned char *w = ws + (size_t)scene * ws_stride;
    float4 *pts = reinterpret_cast<float4 *>(w);                                  // npad x (x, y, z, index)
    float *dist = SDIST ? reinterpret_cast<float *>(dyn) : reinterpret_cast<float *>(w + (size_t)16 * npad);
    unsigned short *cellid = reinterpret_cast<unsigned short *>(w + (size_t)20 * npad);
    // per bucket, in shared memory (after the distances, or -- distances in the workspace -- over the dead histogram):
    // (largest running distance, x, y, z of that candidate) and the candidate's point index
    const int nbcap = npad / BS;
    unsigned char *mbase = dyn + (SDIST ? (size_t)npad * 4 : 0);
    float4 *meta4 = reinterpret_cast<float4 *>(mbase);
    uint32_t *cidx = reinterpret_cast<uint32_t *>(mbase + (size_t)16 * nbcap);

    const size_t nv_off = ((size_t)20 * npad + (size_t)2 * n + 3) & ~(size_t)3;
    int *nv_slot = reinterpret_cast<int *>(w + nv_off);
    int nv = 0;                                          // points that compete (not skipped)
    constexpr int NT = PHASE == 3 ? kSplitThreads : kBT;  // threads of the binning
    constexpr int NWB = NT / 32;                          // ... its warps
    constexpr int SW = PHASE == 3 ? 8 : kBW;              // warps that scan the histogram (the cells divide evenly)
    constexpr int CPW = kCells / SW;
    if constexpr (PHASE == 3) hist = reinterpret_cast<uint32_t *>(w + ((nv_off + 4 + 15) & ~(size_t)15));
    if constexpr (PHASE != 2) {
    for (int i = tid; i < kCells; i += NT) hist[i] = 0u;

    // ---- P1: bounding box of the competing, finite points (+ the contiguous xyz copy) -----------------------
    const bool vec4 = (pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
    for (int k = tid; k < n; k += NT) {
        float x, y, z;
        if (vec4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + (size_t)pitch * k));
            x = v.x; y = v.y; z = v.z;
        } else {
            x = __ldg(p + (size_t)pitch * k); y = __ldg(p + (size_t)pitch * k + 1); z = __ldg(p + (size_t)pitch * k + 2);
        }
        if (cp) { cp[3 * (size_t)k] = x; cp[3 * (size_t)k + 1] = y; cp[3 * (size_t)k + 2] = z; }
        // sampling_gpu.cu:100-101: float mag compared against the double literal 1e-3
        const bool skip = (double)sqnorm3(x, y, z) <= 1e-3;
        if (!skip && finite3(x, y, z)) {
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t lo = __reduce_min_sync(kFullMask, ord_key(mn[a]));
        const uint32_t hi = __reduce_max_sync(kFullMask, ord_key(mx[a]));
        if (lane == 0) { red[warp][a] = lo; red[warp][3 + a] = hi; }
    }
    __syncthreads();
    float lo3[3], ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t lo = __reduce_min_sync(kFullMask, lane < NWB ? red[lane][a] : 0xffffffffu);
        const uint32_t hi = __reduce_max_sync(kFullMask, lane < NWB ? red[lane][3 + a] : 0u);
        lo3[a] = ord_inv(lo);
        ext = fmaxf(ext, ord_inv(hi) - lo3[a]);          // -inf when there is no finite point
    }
    const float scale = (ext > 0.f && ext < INFINITY) ? (float)(1 << kCellBits) / ext : 0.f;
    if (!(fabsf(lo3[0]) < INFINITY)) { lo3[0] = lo3[1] = lo3[2] = 0.f; }

    // ---- P2: Morton cell of every point, histogram ----------------------------------------------------------
    const float *src = cp ? cp : p;
    const int sp = cp ? 3 : pitch;
    constexpr float kQMax = (float)((1 << kCellBits) - 1);
#pragma unroll 4
    for (int k = tid; k < n; k += NT) {
        const float x = src[(size_t)sp * k], y = src[(size_t)sp * k + 1], z = src[(size_t)sp * k + 2];
        uint32_t c = kSkipCell;
        if (!((double)sqnorm3(x, y, z) <= 1e-3)) {
            const uint32_t qx = (uint32_t)fminf(fmaxf((x - lo3[0]) * scale, 0.f), kQMax);
            const uint32_t qy = (uint32_t)fminf(fmaxf((y - lo3[1]) * scale, 0.f), kQMax);
            const uint32_t qz = (uint32_t)fminf(fmaxf((z - lo3[2]) * scale, 0.f), kQMax);
            c = spread5(qx) | (spread5(qy) << 1) | (spread5(qz) << 2);
            atomicAdd(&hist[c], 1u);
        }
        cellid[k] = (unsigned short)c;
    }
    __syncthreads();

    // ---- P3: exclusive scan of the histogram (each warp scans its 2048 cells; warp bases added on use) ------
    if (warp < SW) {
        uint32_t carry = 0;
        const int c0 = warp * CPW;
        for (int i = 0; i < CPW; i += 32) {
            const uint32_t v = hist[c0 + i + lane];
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
                if (lane >= d) inc += t;
            }
            hist[c0 + i + lane] = carry + inc - v;
            carry += __shfl_sync(kFullMask, inc, 31);
        }
        if (lane == 0) red[warp][6] = carry;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < SW ? red[lane][6] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane < SW) wbase[lane] = inc - v;
        if (lane == SW - 1) wbase[SW] = inc;
    }
    __syncthreads();
    nv = (int)wbase[SW];

    // ---- P4: counting sort into the workspace ---------------------------------------------------------------
#pragma unroll 4
    for (int k = tid; k < n; k += NT) {
        const uint32_t c = cellid[k];
        if (c != kSkipCell) {
            const uint32_t pos = atomicAdd(&hist[c], 1u) + wbase[c / CPW];
            pts[pos] = make_float4(src[(size_t)sp * k], src[(size_t)sp * k + 1], src[(size_t)sp * k + 2], __int_as_float(k));
        }
    }
    __syncthreads();                                     // histogram dead from here: `dist` / the bucket records may alias it
    if constexpr (PHASE == 1) {
        if (tid == 0) *nv_slot = nv;
        return;
    }
    } else {
        nv = *nv_slot;                                   // written by the PHASE 1 launch that precedes this one on the stream
    }
    const int NW = nwa;                                  // warps that run the rounds
    if (warp >= NW) return;
    const int nthr = NW * 32;
#define PN2_FPSB_BAR() asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory")

    // ---- P5: running distances, bucket boxes (registers of the owner thread) and first candidates (shared memory) ----
    // super-bucket (ss, l) of warp w = global super-bucket (ss*32 + l)*NW + w: neighbouring super-buckets belong to
    // different warps; its CH buckets are consecutive.  A running distance lives as a float; comparisons of maxima use
    // its bit pattern as a SIGNED integer (non-negative floats order like integers; -1.0 = "no candidate" is negative).
    float slo[SS][3], shi[SS][3], smax[SS];
    float clo[SS][CH][3], chi[SS][CH][3], cmx[SS][CH];
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) {
        smax[ss] = -1.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) { slo[ss][a] = INFINITY; shi[ss][a] = -INFINITY; }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            cmx[ss][c] = -1.f;
#pragma unroll
            for (int a = 0; a < 3; ++a) { clo[ss][c][a] = INFINITY; chi[ss][c][a] = -INFINITY; }
        }
    }
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) {
        for (int l = 0; l < 32; ++l) {
            const long long sup = (long long)(ss * 32 + l) * NW + warp;
            if (sup * CH * BS >= nv) break;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const long long bkt = sup * CH + c, base = bkt * BS;
                if (base >= nv) continue;
                Cand best; best.key = 0u; best.idx = 0u; best.x = best.y = best.z = 0.f;
                float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const long long pos = base + q * 32 + lane;
                    float4 pt = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
                    float d = -1.f;
                    if (pos < nv) { pt = pts[pos]; d = 1e10f; }                     // sampling.cpp:74-76
                    else pts[pos] = pt;                                             // padding: never a candidate
                    dist[pos] = d;
                    if (pos < nv && finite3(pt.x, pt.y, pt.z)) {
                        bl[0] = fminf(bl[0], pt.x); bl[1] = fminf(bl[1], pt.y); bl[2] = fminf(bl[2], pt.z);
                        bh[0] = fmaxf(bh[0], pt.x); bh[1] = fmaxf(bh[1], pt.y); bh[2] = fmaxf(bh[2], pt.z);
                    }
                    const uint32_t key = dist_key(d), pi = (uint32_t)__float_as_int(pt.w);
                    if (cand_better(key, pi, best.key, best.idx, lg_bs, cnt)) { best.key = key; best.idx = pi; best.x = pt.x; best.y = pt.y; best.z = pt.z; }
                }
                const Cand wb = warp_best(best, lg_bs, cnt);
                float rl[3], rh[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    rl[a] = ord_inv(__reduce_min_sync(kFullMask, ord_key(bl[a])));
                    rh[a] = ord_inv(__reduce_max_sync(kFullMask, ord_key(bh[a])));
                }
                if (lane == 0) { meta4[bkt] = make_float4(key_dist(wb.key), wb.x, wb.y, wb.z); cidx[bkt] = wb.idx; }
                if (lane == l) {
                    cmx[ss][c] = key_dist(wb.key);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        clo[ss][c][a] = rl[a]; chi[ss][c][a] = rh[a];
                        slo[ss][a] = fminf(slo[ss][a], rl[a]); shi[ss][a] = fmaxf(shi[ss][a], rh[a]);
                    }
                    smax[ss] = fmaxf(smax[ss], cmx[ss][c]);
                }
            }
        }
    }

    // sampling_gpu.cu:85-87: the first sample is point 0, unconditionally
    const float p0x = __ldg(p), p0y = __ldg(p + 1), p0z = __ldg(p + 2);
    float ox = p0x, oy = p0y, oz = p0z;
    if (tid == 0) {
        out_idx[0] = 0;
        if (out_xyz) { out_xyz[0] = ox; out_xyz[1] = oy; out_xyz[2] = oz; }
    }

    // Work distribution.  Bucket (owner warp w, owner lane l, ss, c) is updated by warp (w + 4c + l) % NW: the dozen
    // buckets around a new sample -- a few neighbouring super-buckets, i.e. neighbouring w, and their children c --
    // land on different warps.  The queue is a bitmap (bit l of word (w, ss, c)); every lane of a consumer warp checks
    // two of its words against a round-invariant pattern of the owner lanes that hash to this warp.
    constexpr int WPL = kWords / 2;                  // bitmap words a consumer lane checks (32 lanes cover kBW * kWords words)
    const int nwords = NW * kWords;
    uint32_t pat[WPL];
    {
        uint32_t p0 = 0u;
        for (int k = 0; k < 32; k += NW) p0 |= 1u << k;
#pragma unroll
        for (int h = 0; h < WPL; ++h) {
            const int x = WPL * lane + h;
            pat[h] = 0u;
            if (x < nwords) {
                const int wsrc = x / kWords, c = x % CH;
                pat[h] = p0 << ((warp + 16 * NW - wsrc - 4 * c) % NW);
            }
        }
    }
    for (int i = tid; i < 2 * kBW * kWords; i += nthr) (&qmask[0][0])[i] = 0u;
    const int kNone = __float_as_int(-1.f);
    // largest (distance bits, then smallest rank of the buckets' candidate points) over the lanes; result in all lanes
    auto better2 = [&](int ak, uint32_t ab, int bk, uint32_t bb) -> bool {
        if (ak != bk) return ak > bk;
        if (ak < 0) return false;
        return rank_of(cidx[ab], lg_bs, cnt) < rank_of(cidx[bb], lg_bs, cnt);
    };
    auto warp_argmax = [&](int &k, uint32_t &bkt) {
        const int kmax = __reduce_max_sync(kFullMask, k);
        const unsigned eqm = __ballot_sync(kFullMask, k == kmax);
        int srcl = __ffs(eqm) - 1;
        if (kmax >= 0 && (eqm & (eqm - 1u))) {
            const uint32_t r = k == kmax ? rank_of(cidx[bkt], lg_bs, cnt) : 0xffffffffu;
            const uint32_t rm = __reduce_min_sync(kFullMask, r);
            srcl = __ffs(__ballot_sync(kFullMask, r == rm)) - 1;
        }
        k = kmax;
        bkt = __shfl_sync(kFullMask, bkt, srcl);
    };
    PN2_FPSB_BAR();

    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0, nproc = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (profiling) tprev = clock64();
    // accumulated over the second half of the rounds only (steady state: a dozen buckets per round; the first rounds
    // touch every bucket and would dominate an average)
#define PN2_FPSB_MARK(i) if (profiling) { const long long tn = clock64(); if (j >= (m >> 1)) pc[i] += tn - tprev; tprev = tn; }

    int cached_k = kNone;                            // this warp's best among its own buckets, valid while none of them is touched
    uint32_t cached_b = 0u;
    bool prev_touched = true;
    struct Buf { float4 p[PPL]; float d[PPL]; float old; };
    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        // ---- 1. which buckets can change?  (owner threads; boxes and bucket maxima in registers) -------------
        uint32_t cm[SS];
#pragma unroll
        for (int ss = 0; ss < SS; ++ss) {
            cm[ss] = 0u;
            if (box_lb(slo[ss][0], slo[ss][1], slo[ss][2], shi[ss][0], shi[ss][1], shi[ss][2], ox, oy, oz) < smax[ss]) {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    if (box_lb(clo[ss][c][0], clo[ss][c][1], clo[ss][c][2], chi[ss][c][0], chi[ss][c][1], chi[ss][c][2], ox, oy, oz) < cmx[ss][c])
                        cm[ss] |= 1u << c;
            }
        }
        // ---- 2. publish them: one ballot per (ss, c) ------------------------------------------------------------
        uint32_t bal[kWords];
        bool touched = false;
#pragma unroll
        for (int ss = 0; ss < SS; ++ss)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                bal[ss * CH + c] = __ballot_sync(kFullMask, (cm[ss] >> c) & 1u);
                touched |= bal[ss * CH + c] != 0u;
            }
        if (lane == 0) {
#pragma unroll
            for (int ss = 0; ss < SS; ++ss)
                *reinterpret_cast<uint4 *>(&qmask[par][warp * kWords + ss * CH]) =
                    make_uint4(bal[ss * CH], bal[ss * CH + 1], bal[ss * CH + 2], bal[ss * CH + 3]);
        }
        PN2_FPSB_MARK(0)
        PN2_FPSB_BAR();
        // ---- 3. every warp picks its share of the queue -------------------------------------------------------
        uint32_t f[WPL];
        bool anyf = false;
        {
            uint32_t q[WPL];
            if (WPL == 2) {
                const uint2 q2 = *reinterpret_cast<const uint2 *>(&qmask[par][2 * lane]);
                q[0] = q2.x; q[1] = q2.y;
            } else {
#pragma unroll
                for (int h4 = 0; h4 < WPL; h4 += 4) {
                    const uint4 q4 = *reinterpret_cast<const uint4 *>(&qmask[par][WPL * lane + h4]);
                    q[h4] = q4.x; q[h4 + 1] = q4.y; q[h4 + 2] = q4.z; q[h4 + 3] = q4.w;
                }
            }
#pragma unroll
            for (int h = 0; h < WPL; ++h) { f[h] = q[h] & pat[h]; anyf |= f[h] != 0u; }    // pat = 0 beyond the active warps' words
        }
        unsigned found = __ballot_sync(kFullMask, anyf);
        if (profiling) { if (found == 0x12345678u) asm volatile("trap;"); nproc += __popc(found); }
        PN2_FPSB_MARK(1)
        int rk = kNone;                              // best of everything this warp accounts for in this round
        uint32_t rb = 0u;
        auto load = [&](uint32_t bkt, Buf &bf) {
            const uint32_t base = bkt * BS;
#pragma unroll
            for (int q = 0; q < PPL; ++q) { bf.p[q] = pts[base + q * 32 + lane]; bf.d[q] = dist[base + q * 32 + lane]; }
            bf.old = meta4[bkt].x;
        };
        auto process = [&](uint32_t bkt, const Buf &bf) {
            const uint32_t base = bkt * BS;
            float nd[PPL];
            bool ch = false;
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                nd[q] = fminf(sqdist3(bf.p[q].x, bf.p[q].y, bf.p[q].z, ox, oy, oz), bf.d[q]);   // sampling_gpu.cu:104-107
                ch |= nd[q] != bf.d[q];
            }
            int key = __float_as_int(bf.old);
            if (__any_sync(kFullMask, ch)) {
                // lane's best point (PPL > 1), then the warp's: the lane that holds it writes the bucket's record
                int bkey = __float_as_int(nd[0]);
                uint32_t bi = (uint32_t)__float_as_int(bf.p[0].w);
                float bx = bf.p[0].x, by = bf.p[0].y, bz = bf.p[0].z;
                if (nd[0] != bf.d[0]) dist[base + lane] = nd[0];
#pragma unroll
                for (int q = 1; q < PPL; ++q) {
                    if (nd[q] != bf.d[q]) dist[base + q * 32 + lane] = nd[q];
                    const int k2 = __float_as_int(nd[q]);
                    const uint32_t pi = (uint32_t)__float_as_int(bf.p[q].w);
                    if (k2 > bkey || (k2 == bkey && k2 >= 0 && rank_of(pi, lg_bs, cnt) < rank_of(bi, lg_bs, cnt))) {
                        bkey = k2; bi = pi; bx = bf.p[q].x; by = bf.p[q].y; bz = bf.p[q].z;
                    }
                }
                key = __reduce_max_sync(kFullMask, bkey);
                const unsigned eqm = __ballot_sync(kFullMask, bkey == key);
                int srcl = __ffs(eqm) - 1;
                if (key >= 0 && (eqm & (eqm - 1u))) {    // several lanes hold the maximum: smallest rank wins
                    const uint32_t r = bkey == key ? rank_of(bi, lg_bs, cnt) : 0xffffffffu;
                    const uint32_t rm = __reduce_min_sync(kFullMask, r);
                    srcl = __ffs(__ballot_sync(kFullMask, r == rm)) - 1;
                }
                if (lane == srcl) { meta4[bkt] = make_float4(__int_as_float(key), bx, by, bz); cidx[bkt] = bi; }
            }
            if (key == rk && key >= 0) __syncwarp();     // the tie-break reads cidx[bkt], possibly just written
            if (better2(key, bkt, rk, rb)) { rk = key; rb = bkt; }
        };
        // entries: for each lane L in `found`, the set bits of its two words; word x = 2L + h -> (owner warp, ss, c)
        uint32_t g[WPL], gl = 0u;
#pragma unroll
        for (int h = 0; h < WPL; ++h) g[h] = 0u;
        auto next = [&](uint32_t &bkt) -> bool {
            bool left = false;
#pragma unroll
            for (int h = 0; h < WPL; ++h) left |= g[h] != 0u;
            if (!left) {
                if (found == 0u) return false;
                gl = (uint32_t)__ffs(found) - 1u; found &= found - 1u;
#pragma unroll
                for (int h = 0; h < WPL; ++h) g[h] = __shfl_sync(kFullMask, f[h], (int)gl);
            }
            uint32_t hh = 0u, bits = 0u;
#pragma unroll
            for (int h = WPL - 1; h >= 0; --h)
                if (g[h]) { hh = (uint32_t)h; bits = g[h]; }
            const uint32_t ol = (uint32_t)__ffs(bits) - 1u;              // owner lane
#pragma unroll
            for (int h = 0; h < WPL; ++h)
                if ((uint32_t)h == hh) g[h] &= g[h] - 1u;
            const uint32_t x = WPL * gl + hh, wsrc = x / kWords, sc = x % kWords;
            bkt = (((sc / CH) * 32u + ol) * (uint32_t)NW + wsrc) * CH + (sc % CH);
            return true;
        };
        Buf A, B;
        uint32_t ba = 0, bb = 0;
        bool ha = next(ba), hb = false;
        if (ha) load(ba, A);
        // while the first loads are in flight: the best of this warp's OWN buckets that are not in the queue (their
        // maxima and candidates do not change this round); cached while none of the warp's buckets is touched
        if (touched || prev_touched) {
            int tk = kNone;
            uint32_t tb = 0u;
#pragma unroll
            for (int ss = 0; ss < SS; ++ss)
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int key = ((cm[ss] >> c) & 1u) ? kNone : __float_as_int(cmx[ss][c]);
                    const uint32_t bkt = (uint32_t)((((ss * 32 + lane) * NW + warp) * CH) + c);
                    if (better2(key, bkt, tk, tb)) { tk = key; tb = bkt; }
                }
            warp_argmax(tk, tb);
            if (!touched) { cached_k = tk; cached_b = tb; }
            rk = tk; rb = tb;
        } else {
            rk = cached_k; rb = cached_b;
        }
        prev_touched = touched;
        while (ha) {
            hb = next(bb);
            if (hb) load(bb, B);
            process(ba, A);
            if (!hb) break;
            ha = next(ba);
            if (ha) load(ba, A);
            process(bb, B);
        }
        if (lane == 0) tableA[warp] = make_int2(rk, (int)rb);
        PN2_FPSB_MARK(2)
        PN2_FPSB_BAR();
        // ---- 4. argmax over the warps' entries (every warp, redundantly) ----------------------------------------
        int bk = kNone;
        uint32_t bbk = 0u;
        if (lane < NW) { const int2 t = tableA[lane]; bk = t.x; bbk = (uint32_t)t.y; }
        if (profiling) { if (bk == 0x7ffffff0) asm volatile("trap;"); }
        PN2_FPSB_MARK(3)
        warp_argmax(bk, bbk);
        const bool none = bk < 0;         // every point skipped: the reference's reduction leaves besti = 0
        uint32_t widx = 0u;
        ox = p0x; oy = p0y; oz = p0z;
        if (!none) {
            const float4 w4 = meta4[bbk];
            ox = w4.y; oy = w4.z; oz = w4.w;
            widx = cidx[bbk];
        }
        if (tid == (j & 127)) {
            out_idx[j] = (int)widx;
            if (out_xyz) { out_xyz[3 * (size_t)j] = ox; out_xyz[3 * (size_t)j + 1] = oy; out_xyz[3 * (size_t)j + 2] = oz; }
        }
        // the owners pick up the new maxima of their updated buckets
#pragma unroll
        for (int ss = 0; ss < SS; ++ss) {
            if (cm[ss]) {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    if ((cm[ss] >> c) & 1u) cmx[ss][c] = meta4[(((ss * 32 + lane) * NW + warp) * CH) + c].x;
                float sx = cmx[ss][0];
#pragma unroll
                for (int c = 1; c < CH; ++c) sx = fmaxf(sx, cmx[ss][c]);
                smax[ss] = sx;
            }
        }
        PN2_FPSB_MARK(4)
    }
    if (profiling) {
        for (int i = 0; i < 6; ++i) prof[i] = pc[i];
        prof[6] = nproc;
    }
#undef PN2_FPSB_MARK
#undef PN2_FPSB_BAR
}

struct BucketCfg {
    int ppl, ss, nwa;
    bool sdist;
    int npad;
    size_t smem, stride;
};

constexpr int kCH = 4;

bool choose(int n, BucketCfg *c)
{
    if (n <= 0) return false;
    int ppl, ss = 1;
    if (n <= kBT * kCH * 32) ppl = 1;
    else if (n <= kBT * kCH * 64) ppl = 2;
    else if (n <= kBT * kCH * 128) ppl = 4;
    else if (n <= 2 * kBT * kCH * 128) { ppl = 4; ss = 2; }
    else return false;
    const int bs = 32 * ppl;
    c->ppl = ppl; c->ss = ss;
    c->npad = (n + bs - 1) / bs * bs;
    const size_t meta = (size_t)20 * (c->npad / bs);              // per bucket: float4 record + candidate index
    // warps that run the rounds: one thread per super-bucket of CH buckets, at least four warps
    const int nsup = (c->npad / bs + kCH - 1) / kCH;
    c->nwa = (nsup + 32 * ss - 1) / (32 * ss);
    if (c->nwa < 4) c->nwa = 4;
    if (c->nwa > kBW) return false;
    c->sdist = (size_t)c->npad * 4 + meta <= kMaxSmem;
    const size_t h = (size_t)kCells * 4, need = (c->sdist ? (size_t)c->npad * 4 : 0) + meta;
    if (need > kMaxSmem) return false;
    c->smem = need > h ? need : h;
    // + the count of competing points (split launch) + the histogram of the two-scenes-per-SM variant
    const size_t bytes = (size_t)20 * c->npad + (size_t)2 * n + 32 + (size_t)kCells * 4;
    c->stride = (bytes + 255) / 256 * 256;
    return true;
}

// Two launches (binning, then the rounds on nwa warps with two scenes per SM): scenes of up to 40 960 points, whose
// rounds need <= kSplitThreads threads; distances live in the workspace (L2).  Opt-in: PN2_FPS_BUCKET_SPLIT=1.
int launch_split(const BucketCfg &c, int b, int n, int m, int lg_bs, int cnt, const float *xyz, int pitch, int *idxs,
                 float *new_xyz, float *xyz_copy, void *ws, long long *prof, cudaStream_t stream)
{
    auto ka = fps_bucket_kernel<1, kCH, 1, false, 1>;
    auto kb = fps_bucket_kernel<1, kCH, 1, false, 2>;
    const size_t hist = (size_t)kCells * 4, meta = (size_t)20 * (c.npad / 32);
    PN2_CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist));
    PN2_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)meta));
    unsigned char *w = static_cast<unsigned char *>(ws);
    ka<<<b, kBT, hist, stream>>>(n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy, w, c.stride, c.npad, c.nwa, prof);
    kb<<<b, c.nwa * 32, meta, stream>>>(n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy, w, c.stride, c.npad, c.nwa, prof);
    count_launches(1);
    PN2_LAUNCH_CHECK("fps_bucket_kernel(split)");
    return PN2_OK;
}

// One launch of kSplitThreads threads, histogram in the workspace: two scenes per SM (PN2_FPS_BUCKET_SPLIT=2).
// Measured at 40 000 points (scripts/gpu_r2_check10.sh): 5.2 ms alone against 4.1 ms; with 296 scenes in one launch
// 0.193 ms of the GPU per 8 scenes against 0.230 ms -- but the pipelined bench step, where these CTAs share the SMs with
// the other kernels of 20-40 batches, runs at 13.7 k scenes/s against 15.4 k.  Off by default.
int launch_compact(const BucketCfg &c, int b, int n, int m, int lg_bs, int cnt, const float *xyz, int pitch, int *idxs,
                   float *new_xyz, float *xyz_copy, void *ws, long long *prof, cudaStream_t stream)
{
    auto kern = fps_bucket_kernel<1, kCH, 1, false, 3>;
    const size_t meta = (size_t)20 * (c.npad / 32);
    PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)meta));
    kern<<<b, kSplitThreads, meta, stream>>>(n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy,
                                             static_cast<unsigned char *>(ws), c.stride, c.npad, c.nwa, prof);
    PN2_LAUNCH_CHECK("fps_bucket_kernel(compact)");
    return PN2_OK;
}

template <int PPL, int SS, bool SDIST>
int launch_cfg(const BucketCfg &c, int b, int n, int m, int lg_bs, int cnt, const float *xyz, int pitch, int *idxs,
               float *new_xyz, float *xyz_copy, void *ws, long long *prof, cudaStream_t stream)
{
    auto kern = fps_bucket_kernel<PPL, kCH, SS, SDIST>;
    PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    kern<<<b, kBT, c.smem, stream>>>(n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy,
                                     static_cast<unsigned char *>(ws), c.stride, c.npad, c.nwa, prof);
    PN2_LAUNCH_CHECK("fps_bucket_kernel");
    return PN2_OK;
}

}  // namespace

// Smallest scene the bucketed kernel is used for, by the caller's objective (PN2_FPS_LATENCY / PN2_FPS_THROUGHPUT).
// Measured on B200 (B = 8, scripts/fps_bucket_bench.py, scripts/gpu_r2_check9.sh):
//   * one launch: at 40 000 points the register-resident cluster kernel is faster (1.36 ms against 4.1 ms: a round here
//     is a chain of ~430 dependent warp instructions); at 200 000 points this kernel is (12.3 ms against 15.6 ms);
//   * SM time, which is what bounds a caller with many batches in flight: this kernel holds ONE SM per scene, 4.1 SM-ms
//     at 40 000 points, the cluster kernel 8 CTAs at three per SM for 2.1 ms = 5.7 SM-ms.  With >= 19 eight-scene
//     batches in flight (148 scenes) the bench step runs at 14.2-15.2 k scenes/s on this kernel against 13.1 k.
// Latency mode: beyond what the cluster kernel covers with 8 CTAs.  Throughput mode: from 32 768 points.
int fps_bucket_min_points(int mode)
{
    const char *e = getenv("PN2_FPS_BUCKET_MIN");    // tests / sweeps: 1 = always, a huge value = never
    if (e) return atoi(e);
    return mode == PN2_FPS_THROUGHPUT ? 32768 : 81921;
}

size_t fps_bucket_workspace_bytes(int b, int n, int mode)
{
    BucketCfg c;
    if (n < fps_bucket_min_points(mode) || !choose(n, &c)) return 0;
    return c.stride * (size_t)b;
}

int fps_bucket_launch(int b, int n, int m, int lg_bs, int cnt, const float *xyz, int pitch, int *idxs, float *new_xyz,
                      float *xyz_copy, void *ws, size_t ws_bytes, long long *prof, cudaStream_t stream)
{
    BucketCfg c;
    if (!choose(n, &c)) return PN2_ERR_INVALID_ARGUMENT;
    if (!ws || ws_bytes < c.stride * (size_t)b || (reinterpret_cast<uintptr_t>(ws) & 15)) return PN2_ERR_WORKSPACE;
#define PN2_FPSB_GO(PPL, SS, SD) \
    return launch_cfg<PPL, SS, SD>(c, b, n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy, ws, prof, stream)
    // Measured on the 20-lane bench step at the driver's 20 steps: 15.1-15.4 k scenes/s as one launch, 14.0 k split
    // (the 512-thread binning CTAs cannot share an SM with two rounds CTAs and queue behind them); 16.6 k against
    // 15.2 k for the split form in a long run with 32 lanes.  Off by default.
    static const int split = [] { const char *e = getenv("PN2_FPS_BUCKET_SPLIT"); return e ? atoi(e) : 0; }();
    if (split == 2 && c.ppl == 1 && c.nwa * 32 <= kSplitThreads)
        return launch_compact(c, b, n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy, ws, prof, stream);
    if (split == 1 && c.ppl == 1 && c.nwa * 32 <= kSplitThreads)
        return launch_split(c, b, n, m, lg_bs, cnt, xyz, pitch, idxs, new_xyz, xyz_copy, ws, prof, stream);
    if (c.ppl == 1) { if (c.sdist) PN2_FPSB_GO(1, 1, true); else PN2_FPSB_GO(1, 1, false); }
    if (c.ppl == 2) PN2_FPSB_GO(2, 1, false);
    if (c.ss == 1) PN2_FPSB_GO(4, 1, false);
    PN2_FPSB_GO(4, 2, false);
#undef PN2_FPSB_GO
}

}  // namespace pn2
