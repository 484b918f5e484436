// query.cu -- ball query and three_nn with the reference's exact index semantics.
//
// ball query (reference: ball_query_gpu.cu:9-44, one block per scene, one thread per centre
// scanning all N points from global memory): the result for a centre is the ascending-index
// prefix (first nsample) of its in-ball set, padded with the first hit; zeros when empty.
// Here a CTA owns 32 centres and splits the N points into S contiguous segments, one per
// warp: lane = centre, warp = segment.  Each warp stages its segment through shared memory
// with coalesced loads (repacked to float4 so the inner loop is one broadcast LDS.128 per
// point) and records, per centre, the first nsample hits of its segment.  Segments are
// index-ordered, so concatenating the per-segment lists in segment order and truncating at
// nsample is exactly the reference's ascending scan.  grid = (ceil(m/32), B): SA1 at B=8 is
// 512 CTAs x 8 warps instead of the reference's 8 blocks.
//
// three_nn (reference: interpolate_gpu.cu:9-59): thread per unknown point, known points
// staged in shared memory; the cascade keeps the reference's strict '<' (earlier index wins
// ties).  The reference keeps its running bests in double initialised to 1e40; comparing a
// float against them is the same as comparing against a float +inf, and (float)1e40 = +inf
// is what it stores when fewer than three known points exist.
#include "common.cuh"
#include <math_constants.h>

namespace pn2 {

constexpr int kBqCentres = 32;
constexpr int kBqTile = 128;          // points staged per warp per step
constexpr int kBqMaxSegments = 8;

__global__ void __launch_bounds__(kBqMaxSegments * 32)
ball_query_kernel(int n, int m, float radius2, int nsample, int nseg, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx)
{
    extern __shared__ __align__(16) unsigned char bq_smem[];
    // layout: tiles[nseg][kBqTile] float4 | counts[nseg][32] int | hits[nseg][nsample][32] int
    float4 *tiles = reinterpret_cast<float4 *>(bq_smem);
    int *counts = reinterpret_cast<int *>(tiles + nseg * kBqTile);
    int *hits = counts + nseg * 32;

    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const size_t bi = blockIdx.y;
    const int c = blockIdx.x * kBqCentres + lane;
    const float *pts = xyz + bi * (size_t)n * 3;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (c < m) {
        const float *q = new_xyz + (bi * m + c) * 3;
        cx = __ldg(q); cy = __ldg(q + 1); cz = __ldg(q + 2);
    }
    // segment bounds: multiples of the tile so that staging stays aligned
    const int per = ((n + nseg - 1) / nseg + kBqTile - 1) / kBqTile * kBqTile;
    const int k0 = min(n, seg * per), k1 = min(n, k0 + per);
    float4 *tile = tiles + seg * kBqTile;
    int *myhits = hits + (size_t)seg * nsample * 32 + lane;
    int cnt = c < m ? 0 : nsample;   // out-of-range lanes are "full" from the start

    for (int base = k0; base < k1; base += kBqTile) {
        if (__all_sync(0xffffffffu, cnt >= nsample)) break;
        const int len = min(kBqTile, k1 - base);
        __syncwarp();
        for (int t = lane; t < len; t += 32) {
            const float *s = pts + (size_t)(base + t) * 3;
            tile[t] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
        }
        __syncwarp();
#pragma unroll 4
        for (int t = 0; t < len; ++t) {
            const float4 pt = tile[t];
            // ball_query_gpu.cu:31-33: (new - x)^2 terms, contracted FMUL,FFMA,FFMA; strict '<'
            const float d2 = sqdist3(cx, cy, cz, pt.x, pt.y, pt.z);
            if (d2 < radius2 && cnt < nsample) {
                myhits[cnt * 32] = base + t;
                ++cnt;
            }
        }
    }
    if (c >= m) cnt = 0;
    counts[seg * 32 + lane] = cnt;
    __syncthreads();

    // merge: warp w writes the rows of centres w, w + nseg, ...
    for (int cc = seg; cc < kBqCentres; cc += nseg) {
        const int centre = blockIdx.x * kBqCentres + cc;
        if (centre >= m) break;
        int total = 0, first = 0;
        bool have_first = false;
        for (int s = 0; s < nseg; ++s) {
            const int v = counts[s * 32 + cc];
            if (!have_first && v > 0) { first = hits[(size_t)s * nsample * 32 + cc]; have_first = true; }
            total += v;
        }
        int *out = idx + (bi * m + centre) * (size_t)nsample;
        for (int l = lane; l < nsample; l += 32) {
            int v = first;   // padding value (0 when the ball is empty: ball_query.cpp:19-21)
            if (l < total) {
                int rem = l;
                for (int s = 0; s < nseg; ++s) {
                    const int cs = counts[s * 32 + cc];
                    if (rem < cs) { v = hits[((size_t)s * nsample + rem) * 32 + cc]; break; }
                    rem -= cs;
                }
            }
            out[l] = v;
        }
    }
}

constexpr int kNnThreads = 128;
constexpr int kNnTile = 512;

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx)
{
    __shared__ float4 tile[kNnTile];
    const size_t bi = blockIdx.y;
    const int j = blockIdx.x * kNnThreads + threadIdx.x;
    const float *kn = known + bi * (size_t)m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (j < n) {
        const float *u = unknown + (bi * n + j) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += kNnTile) {
        const int len = min(kNnTile, m - base);
        __syncthreads();
        for (int t = threadIdx.x; t < len; t += kNnThreads) {
            const float *s = kn + (size_t)(base + t) * 3;
            tile[t] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
        }
        __syncthreads();
        for (int t = 0; t < len; ++t) {
            const float4 pt = tile[t];
            const float d = sqdist3_yxz(ux, uy, uz, pt.x, pt.y, pt.z);   // interpolate_gpu.cu:33-34
            const int k = base + t;
            if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
            else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
            else if (d < b3) { b3 = d; i3 = k; }
        }
    }
    if (j < n) {
        float *od = dist2 + (bi * n + j) * 3;
        int *oi = idx + (bi * n + j) * 3;
        od[0] = b1; od[1] = b2; od[2] = b3;
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, pn2_stream_t stream)
{
    if (b < 0 || n < 0 || m < 0 || nsample < 0 || b > 65535) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || m == 0 || nsample == 0) return PN2_OK;
    if (!new_xyz || !idx || (n > 0 && !xyz)) return PN2_ERR_INVALID_ARGUMENT;
    if (n == 0) {
        PN2_CUDA_TRY(cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)b * m * nsample, as_stream(stream)));
        return PN2_OK;
    }
    // as many segments as fit in shared memory, at most 8 and not more than the scene has tiles
    const size_t per_seg = sizeof(float4) * kBqTile + sizeof(int) * 32 + sizeof(int) * 32 * (size_t)nsample;
    int nseg = (int)min((size_t)kBqMaxSegments, (size_t)(200 * 1024) / per_seg);
    nseg = min(nseg, max(1, ceil_div(n, kBqTile)));
    if (nseg < 1) return PN2_ERR_INVALID_ARGUMENT;   // nsample > ~1500: one segment's hit list exceeds shared memory
    const size_t smem = per_seg * nseg;
    static bool attr_set = false;
    if (!attr_set) {
        PN2_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          220 * 1024));
        attr_set = true;
    }
    const float radius2 = radius * radius;   // ball_query_gpu.cu:22, fp32 product
    dim3 grid(ceil_div(m, kBqCentres), b);
    ball_query_kernel<<<grid, nseg * 32, smem, as_stream(stream)>>>(n, m, radius2, nsample, nseg, new_xyz,
                                                                     xyz, idx);
    PN2_LAUNCH_CHECK("ball_query");
    return PN2_OK;
}

extern "C" int pn2_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                            int *idx, pn2_stream_t stream)
{
    if (b < 0 || n < 0 || m < 0 || b > 65535) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || n == 0) return PN2_OK;
    if (!unknown || !dist2 || !idx || (m > 0 && !known)) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(n, kNnThreads), b);
    three_nn_kernel<<<grid, kNnThreads, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
    PN2_LAUNCH_CHECK("three_nn");
    return PN2_OK;
}
