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

// ---- uniform-grid ball query --------------------------------------------------------------------
// For large scenes (SA1: 2048 centres x 40 000 points) the scan above tests 82 M pairs per scene although
// a ball holds ~50 points.  The grid variant bins the points into cells of edge >= 1.001 * radius
// (<= 40 cells per axis), sorts them by cell (count / scan / scatter) and lets ONE WARP per centre test
// only the 27 surrounding cells with coalesced float4 loads of the cell-sorted copy.
// Exactness: (i) the hit test is the same sqdist3 < r*r on the same floats, so the hit SET is the
// reference's; a pair closer than r differs by less than one cell per axis (cells are 0.1 % larger than r
// and the cell quotient stays below 2^11, so float rounding of the quotient cannot bridge the margin);
// (ii) the reference returns the hits in ascending original index: every hit sets one bit of a per-warp
// bitmap over the point indices, each lane owns a contiguous slice of the bitmap, a warp prefix sum over
// the slices' popcounts gives each lane its output offset, and the lanes emit their set bits in order.
constexpr int kGridMaxAxis = 40;
constexpr int kGridMaxCells = kGridMaxAxis * kGridMaxAxis * kGridMaxAxis;
constexpr int kGridMinPoints = 4096;      // below this the plain scan is faster than building a grid

struct GridParams {        // per scene, written by bq_grid_setup_kernel
    float minx, miny, minz;
    float invx, invy, invz;
    int nx, ny, nz, ncells;
};

__device__ __forceinline__ int grid_axis_cell(float v, float mn, float inv, int n)
{
    // unclamped cell of a coordinate, limited to [-2, n+1] so that the int conversion is defined
    const float q = fminf(fmaxf((v - mn) * inv, -2.f), (float)(n + 1));
    return (int)floorf(q);
}

__global__ void __launch_bounds__(1024)
bq_grid_setup_kernel(int n, int pitch, float cell_min, const float *__restrict__ xyz,
                     GridParams *__restrict__ params)
{
    __shared__ float red[6][32];
    const size_t bi = blockIdx.x;
    const float *p = xyz + bi * (size_t)n * pitch;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = threadIdx.x; k < n; k += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(p + (size_t)pitch * k + a);
            if (isfinite(v)) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int a = 0; a < 3; ++a) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int a = 0; a < 3; ++a)
            for (int w = 1; w < nw; ++w) { red[a][0] = fminf(red[a][0], red[a][w]); red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]); }
        GridParams g;
        float mn[3], inv[3];
        int dim[3];
        for (int a = 0; a < 3; ++a) {
            mn[a] = red[a][0];
            const float ext = fmaxf(red[3 + a][0] - red[a][0], 0.f);
            const float cell = fmaxf(cell_min, ext / (float)kGridMaxAxis);
            inv[a] = 1.0f / cell;
            dim[a] = min(kGridMaxAxis, (int)floorf(ext * inv[a]) + 1);
            if (!(ext >= 0.f) || !isfinite(ext)) { mn[a] = 0.f; dim[a] = 1; }   // no finite coordinate on this axis
        }
        g.minx = mn[0]; g.miny = mn[1]; g.minz = mn[2];
        g.invx = inv[0]; g.invy = inv[1]; g.invz = inv[2];
        g.nx = dim[0]; g.ny = dim[1]; g.nz = dim[2];
        g.ncells = dim[0] * dim[1] * dim[2];
        params[bi] = g;
    }
}

__device__ __forceinline__ int grid_point_cell(const GridParams &g, float x, float y, float z)
{
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return -1;   // can never be a hit (NaN / inf distance)
    const int cx = min(max(grid_axis_cell(x, g.minx, g.invx, g.nx), 0), g.nx - 1);
    const int cy = min(max(grid_axis_cell(y, g.miny, g.invy, g.ny), 0), g.ny - 1);
    const int cz = min(max(grid_axis_cell(z, g.minz, g.invz, g.nz), 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}

// pass 0: count points per cell; pass 1: scatter (x, y, z, index) into cell order
__global__ void __launch_bounds__(256)
bq_grid_bin_kernel(int n, int pitch, int pass, const float *__restrict__ xyz,
                   const GridParams *__restrict__ params, int *__restrict__ cursor, float4 *__restrict__ sorted)
{
    const size_t bi = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const GridParams g = params[bi];
    const float *p = xyz + (bi * n + k) * (size_t)pitch;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const int cell = grid_point_cell(g, x, y, z);
    if (cell < 0) return;
    int *cur = cursor + bi * (size_t)(kGridMaxCells + 1);
    if (pass == 0) atomicAdd(cur + cell, 1);
    else sorted[bi * n + atomicAdd(cur + cell, 1)] = make_float4(x, y, z, __int_as_float(k));
}

// exclusive scan of the cell counts of one scene: start[c] and cursor[c] = first slot of cell c
__global__ void __launch_bounds__(1024)
bq_grid_scan_kernel(const GridParams *__restrict__ params, int *__restrict__ cursor, int *__restrict__ start)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const size_t bi = blockIdx.x;
    const int ncells = params[bi].ncells;
    int *cur = cursor + bi * (size_t)(kGridMaxCells + 1);
    int *st = start + bi * (size_t)(kGridMaxCells + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += 1024) {
        const int c = base + threadIdx.x;
        const int v = c < ncells ? cur[c] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - v;
        if (c < ncells) { st[c] = excl; cur[c] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) st[ncells] = carry;
}

constexpr int kGqWarps = 8;

__global__ void __launch_bounds__(kGqWarps * 32)
bq_grid_query_kernel(int n, int m, float radius2, int nsample, int wpl, int warps, const float *__restrict__ new_xyz,
                     const GridParams *__restrict__ params, const int *__restrict__ start,
                     const float4 *__restrict__ sorted, int *__restrict__ idx)
{
    extern __shared__ __align__(16) unsigned int gq_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t bi = blockIdx.y;
    const int c = blockIdx.x * warps + warp;
    const int words = wpl * 32;
    unsigned int *bm = gq_smem + (size_t)warp * (words + nsample);   // this warp's bitmap, then its output row
    int *stage = reinterpret_cast<int *>(bm + words);
    for (int i = lane; i < words; i += 32) bm[i] = 0u;
    __syncwarp();
    if (c >= m) return;
    const GridParams g = params[bi];
    const float *q = new_xyz + (bi * m + c) * 3;
    const float cx = __ldg(q), cy = __ldg(q + 1), cz = __ldg(q + 2);
    const int *st = start + bi * (size_t)(kGridMaxCells + 1);
    const float4 *pts = sorted + bi * n;
    if (isfinite(cx) && isfinite(cy) && isfinite(cz)) {
        const int gx = grid_axis_cell(cx, g.minx, g.invx, g.nx);
        const int gy = grid_axis_cell(cy, g.miny, g.invy, g.ny);
        const int gz = grid_axis_cell(cz, g.minz, g.invz, g.nz);
        const int x0 = max(gx - 1, 0), x1 = min(gx + 1, g.nx - 1);
        if (x0 <= x1)
            for (int zz = max(gz - 1, 0); zz <= min(gz + 1, g.nz - 1); ++zz)
                for (int yy = max(gy - 1, 0); yy <= min(gy + 1, g.ny - 1); ++yy) {
                    const int row = (zz * g.ny + yy) * g.nx;
                    const int s = __ldg(st + row + x0), e = __ldg(st + row + x1 + 1);   // x-adjacent cells are contiguous
                    for (int pos = s + lane; pos < e; pos += 32) {
                        const float4 pt = __ldg(pts + pos);
                        // ball_query_gpu.cu:31-33: same expression, same operand order, strict '<'
                        if (sqdist3(cx, cy, cz, pt.x, pt.y, pt.z) < radius2) {
                            const int k = __float_as_int(pt.w);
                            atomicOr(bm + (k >> 5), 1u << (k & 31));
                        }
                    }
                }
    }
    __syncwarp();
    // lane l owns bitmap words [l*wpl, (l+1)*wpl): count, prefix over lanes, emit in ascending index order
    unsigned int *mine = bm + lane * wpl;
    int cnt = 0;
    for (int i = 0; i < wpl; ++i) cnt += __popc(mine[i]);
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - cnt;
    for (int i = 0; i < wpl && pos < nsample && cnt > 0; ++i) {
        unsigned int w = mine[i];
        while (w != 0u && pos < nsample) {
            const int b = __ffs(w) - 1;
            stage[pos++] = (lane * wpl + i) * 32 + b;
            w &= w - 1u;
        }
    }
    __syncwarp();
    const int first = total > 0 ? stage[0] : 0;            // padding: the first hit; zeros for an empty ball
    int *out = idx + (bi * m + c) * (size_t)nsample;
    for (int l = lane; l < nsample; l += 32) out[l] = l < total ? stage[l] : first;
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
    // per launch: the attribute belongs to the current device (one process may drive several GPUs)
    PN2_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const float radius2 = radius * radius;   // ball_query_gpu.cu:22, fp32 product
    dim3 grid(ceil_div(m, kBqCentres), b);
    ball_query_kernel<<<grid, nseg * 32, smem, as_stream(stream)>>>(n, m, radius2, nsample, nseg, new_xyz,
                                                                     xyz, idx);
    PN2_LAUNCH_CHECK("ball_query");
    return PN2_OK;
}

extern "C" size_t pn2_ball_query_workspace_bytes(int b, int n, int m, int nsample)
{
    (void)m; (void)nsample;
    if (b <= 0 || n < kGridMinPoints) return 0;
    return (size_t)b * (sizeof(GridParams) + 2 * sizeof(int) * (size_t)(kGridMaxCells + 1) + sizeof(float4) * (size_t)n) + 256;
}

namespace {

struct GridWs {
    GridParams *params;
    float4 *sorted;
    int *cursor, *start;
    int wpl, warps;
    size_t per_warp;
};

// Carves the workspace; false when this shape takes the plain scan (small scene, no/short workspace, a ball
// that cannot be realised as a grid).
bool grid_ws(int b, int n, int m, float radius, int nsample, void *workspace, size_t workspace_bytes, GridWs &g)
{
    const size_t need = pn2_ball_query_workspace_bytes(b, n, m, nsample);
    g.wpl = (ceil_div(n, 1024)) | 1;                             // bitmap words per lane, odd: conflict-free slices
    g.per_warp = sizeof(int) * ((size_t)g.wpl * 32 + nsample);
    g.warps = (int)min((size_t)kGqWarps, (size_t)(200 * 1024) / max(g.per_warp, (size_t)1));
    if (need == 0 || !workspace || workspace_bytes < need || g.warps < 1 || !(radius > 0.f) || m <= 0 || nsample <= 0)
        return false;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    w += (256 - (reinterpret_cast<uintptr_t>(w) & 255)) & 255;
    g.params = reinterpret_cast<GridParams *>(w);
    g.sorted = reinterpret_cast<float4 *>(w + (((size_t)b * sizeof(GridParams) + 15) & ~(size_t)15));
    g.cursor = reinterpret_cast<int *>(g.sorted + (size_t)b * n);
    g.start = g.cursor + (size_t)b * (kGridMaxCells + 1);
    return true;
}

}   // namespace

// The grid depends on the points and the radius only, not on the centres: a caller whose centres come from a
// sampling kernel on another stream builds the grid while that kernel runs and queries it afterwards.
extern "C" int pn2_ball_query_grid_build(int b, int n, int m, float radius, int nsample, const float *xyz,
                                         int xyz_pitch, void *workspace, size_t workspace_bytes,
                                         pn2_stream_t stream)
{
    GridWs g;
    if (b <= 0 || n <= 0 || b > 65535 || !xyz || xyz_pitch < 3) return PN2_ERR_INVALID_ARGUMENT;
    if (!grid_ws(b, n, m, radius, nsample, workspace, workspace_bytes, g)) return PN2_ERR_WORKSPACE;
    cudaStream_t s = as_stream(stream);
    PN2_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, sizeof(int) * (size_t)b * (kGridMaxCells + 1), s));
    bq_grid_setup_kernel<<<b, 1024, 0, s>>>(n, xyz_pitch, radius * 1.001f, xyz, g.params);
    dim3 pgrid(ceil_div(n, 256), b);
    bq_grid_bin_kernel<<<pgrid, 256, 0, s>>>(n, xyz_pitch, 0, xyz, g.params, g.cursor, g.sorted);
    bq_grid_scan_kernel<<<b, 1024, 0, s>>>(g.params, g.cursor, g.start);
    bq_grid_bin_kernel<<<pgrid, 256, 0, s>>>(n, xyz_pitch, 1, xyz, g.params, g.cursor, g.sorted);
    count_launches(3);   // setup, bin x2, scan (one is counted by the check below)
    PN2_LAUNCH_CHECK("ball_query_grid_build");
    return PN2_OK;
}

extern "C" int pn2_ball_query_grid_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                         int *idx, void *workspace, size_t workspace_bytes, pn2_stream_t stream)
{
    GridWs g;
    if (b <= 0 || n <= 0 || b > 65535 || !new_xyz || !idx) return PN2_ERR_INVALID_ARGUMENT;
    if (!grid_ws(b, n, m, radius, nsample, workspace, workspace_bytes, g)) return PN2_ERR_WORKSPACE;
    PN2_CUDA_TRY(cudaFuncSetAttribute(bq_grid_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    dim3 qgrid(ceil_div(m, g.warps), b);
    bq_grid_query_kernel<<<qgrid, g.warps * 32, g.per_warp * g.warps, as_stream(stream)>>>(
        n, m, radius * radius, nsample, g.wpl, g.warps, new_xyz, g.params, g.start, g.sorted, idx);
    PN2_LAUNCH_CHECK("ball_query_grid");
    return PN2_OK;
}

extern "C" int pn2_ball_query_ws(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                 const float *xyz, int *idx, void *workspace, size_t workspace_bytes,
                                 pn2_stream_t stream)
{
    GridWs g;
    // small scenes, balls that cannot be realised as a grid, or no workspace: the plain scan
    if (!grid_ws(b, n, m, radius, nsample, workspace, workspace_bytes, g))
        return pn2_ball_query(b, n, m, radius, nsample, new_xyz, xyz, idx, stream);
    if (b > 65535 || !new_xyz || !xyz || !idx) return PN2_ERR_INVALID_ARGUMENT;
    const int rc = pn2_ball_query_grid_build(b, n, m, radius, nsample, xyz, 3, workspace, workspace_bytes, stream);
    if (rc != PN2_OK) return rc;
    return pn2_ball_query_grid_query(b, n, m, radius, nsample, new_xyz, idx, workspace, workspace_bytes, stream);
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
