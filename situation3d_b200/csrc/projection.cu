// projection.cu -- 3-D point <-> image pixel correspondences and feature back-projection for the multiview
// features.
//
// Reference: lib/projection.py, class ProjectionHelper.
//   compute_projection (:191-254): per camera view, ~25 PyTorch launches plus three host synchronisations
//     (`mask.any()`): frustum corners (bmm, :68), six plane normals (torch.cross, :85-117), the half-space test
//     round(dot * 100) / 100 < 0 per plane (:143-146), world -> camera (mm with the inverse pose, :223), pinhole
//     projection and round-to-pixel (:226-228), image-range test (:231), depth look-up and the three depth tests
//     (:239-240), then compaction into two (num_points + 1) int64 arrays: element 0 = number of correspondences,
//     then the point indices (ascending) / the pixel indices y * width + x, zero-padded (:246-252).
//   project (:257-279): output (C, num_points) zero-filled, output[:, indices_3d] = label[:, indices_2d].
// Here all views of a scene go through two launches (flags + per-segment counts; ordered compaction + zero fill),
// and the back-projection is one dense pass over the output (a per-point pixel map replaces the scatter, so the
// zero fill and the gather are the same coalesced store).
//
// Arithmetic: the reference leaves the summation order of its 3- and 4-term dot products to cuBLAS; here every
// expression is written out (FMA chains in index order for the matrix products, separately rounded products for
// the cross products and element-wise steps, IEEE division) and oracle/pn2_oracle.c restates it with fmaf, so
// kernel and oracle agree bit for bit.  Against the reference itself the index lists can differ only for a point
// whose plane distance or pixel coordinate sits within an ulp of a rounding boundary.
#include "common.cuh"

namespace pn2 {

constexpr int kProjThreads = 256;
constexpr int kProjSegment = 4096;            // points per CTA: 16 per thread

struct ProjCamera {                            // pn2_compute_projection's scalar arguments
    float fx, fy, cx, cy;                      // intrinsic[0][0], [1][1], [0][2], [1][2]
    float depth_min, depth_max, accuracy;
    int width, height;                         // image_dims[0], image_dims[1]
    float corner[8][3];                        // ProjectionHelper._compute_corner_points (camera frame, w = 1)
};

struct ProjPlanes {
    float c2[3], c4[3];                        // the two corners the half-space tests are anchored on (:139-140)
    float normal[6][3];
};

// corner_coords = camera_to_world @ [corner, 1] (:68); normals as :85-117.  `corners` (8 x 4) may be NULL.
__device__ __forceinline__ void frustum_planes(const float *__restrict__ m, const ProjCamera &cam, ProjPlanes &pl, float *corners)
{
    float c[8][3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float acc = __fmul_rn(m[r * 4 + 0], cam.corner[k][0]);
            acc = __fmaf_rn(m[r * 4 + 1], cam.corner[k][1], acc);
            acc = __fmaf_rn(m[r * 4 + 2], cam.corner[k][2], acc);
            acc = __fmaf_rn(m[r * 4 + 3], 1.0f, acc);
            if (r < 3) c[k][r] = acc;
            if (corners) corners[k * 4 + r] = acc;
        }
    }
    const int tri[6][3] = {{0, 3, 1}, {1, 2, 5}, {2, 3, 6}, {3, 0, 7}, {0, 1, 4}, {5, 6, 4}};   // origin, end of vec1, end of vec2
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float a[3], b[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = __fsub_rn(c[tri[k][1]][j], c[tri[k][0]][j]);
            b[j] = __fsub_rn(c[tri[k][2]][j], c[tri[k][0]][j]);
        }
        pl.normal[k][0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
        pl.normal[k][1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
        pl.normal[k][2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) { pl.c2[j] = c[2][j]; pl.c4[j] = c[4][j]; }
}

// points_in_frustum (:121-155): planes 0-2 are measured from corner 2, planes 3-5 from corner 4.
__device__ __forceinline__ bool point_in_frustum(float x, float y, float z, const ProjPlanes &pl)
{
    const float ax = __fsub_rn(x, pl.c2[0]), ay = __fsub_rn(y, pl.c2[1]), az = __fsub_rn(z, pl.c2[2]);
    const float bx = __fsub_rn(x, pl.c4[0]), by = __fsub_rn(y, pl.c4[1]), bz = __fsub_rn(z, pl.c4[2]);
    bool in = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float px = k < 3 ? ax : bx, py = k < 3 ? ay : by, pz = k < 3 ? az : bz;
        const float d = __fmaf_rn(pz, pl.normal[k][2], __fmaf_rn(py, pl.normal[k][1], __fmul_rn(px, pl.normal[k][0])));
        in = in && (rintf(__fmul_rn(d, 100.0f)) < 0.0f);           // round(d * 100) / 100 < 0
    }
    return in;
}

// The pixel index y * width + x a point corresponds to in this view, or -1 (:214-240 as one predicate).
__device__ __forceinline__ int project_point(float x, float y, float z, const ProjPlanes &pl, const float *__restrict__ w2c,
                                             const float *__restrict__ depth, const ProjCamera &cam)
{
    if (!point_in_frustum(x, y, z, pl)) return -1;
    float cam3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(w2c[r * 4 + 0], x);
        acc = __fmaf_rn(w2c[r * 4 + 1], y, acc);
        acc = __fmaf_rn(w2c[r * 4 + 2], z, acc);
        cam3[r] = __fmaf_rn(w2c[r * 4 + 3], 1.0f, acc);
    }
    const float u = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cam3[0], cam.fx), cam3[2]), cam.cx));
    const float v = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cam3[1], cam.fy), cam3[2]), cam.cy));
    if (!(u >= 0.0f && v >= 0.0f && u < (float)cam.width && v < (float)cam.height)) return -1;    // NaN / inf fail here
    const int pix = (int)v * cam.width + (int)u;
    const float dv = depth[pix];
    if (!(dv >= cam.depth_min && dv <= cam.depth_max && fabsf(__fsub_rn(dv, cam3[2])) <= cam.accuracy)) return -1;
    return pix;
}

__global__ void __launch_bounds__(32)
frustum_planes_kernel(int views, const float *__restrict__ c2w, ProjCamera cam, float *__restrict__ corners, float *__restrict__ normals)
{
    const int v = blockIdx.x * 32 + threadIdx.x;
    if (v >= views) return;
    ProjPlanes pl;
    frustum_planes(c2w + (size_t)v * 16, cam, pl, corners ? corners + (size_t)v * 32 : nullptr);
    if (normals)
        for (int k = 0; k < 6; ++k)
            for (int j = 0; j < 3; ++j) normals[(size_t)v * 18 + k * 3 + j] = pl.normal[k][j];
}

// points_in_frustum as the reference exposes it: the caller's corners (8,4) and normals (6,3), a byte mask and a count.
__global__ void __launch_bounds__(kProjThreads)
points_in_frustum_kernel(int n, const float *__restrict__ points, const float *__restrict__ corners, const float *__restrict__ normals,
                         unsigned char *__restrict__ mask, int *__restrict__ count)
{
    __shared__ ProjPlanes pl;
    if (threadIdx.x < 18) pl.normal[threadIdx.x / 3][threadIdx.x % 3] = normals[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 35) { pl.c2[threadIdx.x - 32] = corners[2 * 4 + threadIdx.x - 32]; pl.c4[threadIdx.x - 32] = corners[4 * 4 + threadIdx.x - 32]; }
    __syncthreads();
    int cnt = 0;
    for (int i = blockIdx.x * kProjThreads + threadIdx.x; i < n; i += gridDim.x * kProjThreads) {
        const bool in = point_in_frustum(points[(size_t)i * 3], points[(size_t)i * 3 + 1], points[(size_t)i * 3 + 2], pl);
        if (mask) mask[i] = in ? 1 : 0;
        cnt += in;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(count, cnt);
}

// Pass 1: pixel-or-minus-one per (view, point) and the number of correspondences per (view, segment).
__global__ void __launch_bounds__(kProjThreads)
projection_flags_kernel(int n, int nseg, const float *__restrict__ points, const float *__restrict__ depth,
                        const float *__restrict__ c2w, const float *__restrict__ w2c, ProjCamera cam,
                        int *__restrict__ pixmap, int *__restrict__ segcount)
{
    __shared__ ProjPlanes pl;
    __shared__ float w[12];
    __shared__ int warp_cnt[kProjThreads / 32];
    const int view = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) frustum_planes(c2w + (size_t)view * 16, cam, pl, nullptr);
    if (tid >= 32 && tid < 44) w[tid - 32] = w2c[(size_t)view * 16 + tid - 32];
    __syncthreads();
    const float *dmap = depth + (size_t)view * cam.width * cam.height;
    const int lo = seg * kProjSegment, hi = min(lo + kProjSegment, n);
    int cnt = 0;
    for (int i = lo + tid; i < hi; i += kProjThreads) {
        const int pix = project_point(points[(size_t)i * 3], points[(size_t)i * 3 + 1], points[(size_t)i * 3 + 2], pl, w, dmap, cam);
        pixmap[(size_t)view * n + i] = pix;
        cnt += pix >= 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int k = 0; k < kProjThreads / 32; ++k) s += warp_cnt[k];
        segcount[(size_t)view * nseg + seg] = s;
    }
}

// Pass 2: ordered compaction of a segment to its place in the two lists, zero fill of the unused tail (:246-252).
__global__ void __launch_bounds__(kProjThreads)
projection_compact_kernel(int n, int nseg, const int *__restrict__ pixmap, const int *__restrict__ segcount,
                          long long *__restrict__ idx3d, long long *__restrict__ idx2d, int *__restrict__ counts)
{
    __shared__ int warp_off[kProjThreads / 32 + 1];
    __shared__ int s_base, s_total;
    const int view = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (wid == 0) {
        int before = 0, all = 0;
        for (int k = lane; k < nseg; k += 32) {
            const int c = segcount[(size_t)view * nseg + k];
            all += c;
            before += k < seg ? c : 0;
        }
        before = __reduce_add_sync(0xffffffffu, before);
        all = __reduce_add_sync(0xffffffffu, all);
        if (lane == 0) { s_base = before; s_total = all; }
    }
    __syncthreads();
    const int total = s_total;
    long long *o3 = idx3d + (size_t)view * (n + 1), *o2 = idx2d + (size_t)view * (n + 1);
    if (seg == 0 && tid == 0) {
        o3[0] = total;
        o2[0] = total;
        if (counts) counts[view] = total;
    }
    const int lo = seg * kProjSegment, hi = min(lo + kProjSegment, n);
    int base = s_base;
    for (int i0 = lo; i0 < hi; i0 += kProjThreads) {
        const int i = i0 + tid;
        const int pix = i < hi ? pixmap[(size_t)view * n + i] : -1;
        const unsigned ballot = __ballot_sync(0xffffffffu, pix >= 0);
        if (lane == 0) warp_off[wid + 1] = __popc(ballot);
        __syncthreads();
        if (tid == 0) {
            warp_off[0] = 0;
            for (int k = 1; k <= kProjThreads / 32; ++k) warp_off[k] += warp_off[k - 1];
        }
        __syncthreads();
        if (pix >= 0) {
            const int slot = base + warp_off[wid] + __popc(ballot & ((1u << lane) - 1u));
            o3[1 + slot] = i;
            o2[1 + slot] = pix;
        }
        base += warp_off[kProjThreads / 32];
        __syncthreads();
        // slot i of the lists is never written by a correspondence when i >= total: this CTA zero-fills its share
        if (i < hi && i >= total) { o3[1 + i] = 0; o2[1 + i] = 0; }
    }
}

// project(): the per-point pixel map of one view from its two lists ...
__global__ void __launch_bounds__(kProjThreads)
project_map_kernel(int n, int hw, const long long *__restrict__ idx3d, const long long *__restrict__ idx2d, int *__restrict__ pixmap,
                   int *__restrict__ status)
{
    const int view = blockIdx.y;
    const long long *i3 = idx3d + (size_t)view * (n + 1), *i2 = idx2d + (size_t)view * (n + 1);
    const long long cnt = i3[0];
    if (cnt < 0 || cnt > n) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(status, 1); return; }
    for (long long k = (long long)blockIdx.x * kProjThreads + threadIdx.x; k < cnt; k += (long long)gridDim.x * kProjThreads) {
        const long long pt = i3[1 + k], pix = i2[1 + k];
        if (pt < 0 || pt >= n || pix < 0 || pix >= hw) { atomicExch(status, 1); continue; }
        pixmap[(size_t)view * n + pt] = (int)pix;
    }
}

// ... and the dense pass: out[view, c, pt] = label[view, c, pixmap[view, pt]] or 0.  One thread per point, a group of
// channels per CTA row; the stores of a warp are 128 contiguous bytes, the label plane of a view stays in L2.
template <int CPB>
__global__ void __launch_bounds__(kProjThreads)
project_dense_kernel(int n, int c, int hw, const int *__restrict__ pixmap, const float *__restrict__ label, float *__restrict__ out)
{
    const int view = blockIdx.z, c0 = blockIdx.y * CPB;
    const int pt = blockIdx.x * kProjThreads + threadIdx.x;
    if (pt >= n) return;
    const int pix = pixmap[(size_t)view * n + pt];
    const float *lab = label + ((size_t)view * c + c0) * hw;
    float *o = out + ((size_t)view * c + c0) * n + pt;
    const int nc = min(CPB, c - c0);
    if (nc == CPB) {
#pragma unroll
        for (int j0 = 0; j0 < CPB; j0 += 8) {
            float vals[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) vals[j] = pix >= 0 ? __ldg(lab + (size_t)(j0 + j) * hw + pix) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) __stcs(o + (size_t)(j0 + j) * n, vals[j]);
        }
    } else {
        for (int j = 0; j < nc; ++j) __stcs(o + (size_t)j * n, pix >= 0 ? __ldg(lab + (size_t)j * hw + pix) : 0.f);
    }
}

static bool fill_camera(ProjCamera &cam, const float *intrinsic4, const float *depth_range3, int width, int height,
                        const float *corner_points)
{
    if (!intrinsic4 || !depth_range3 || !corner_points || width < 1 || height < 1 || (long long)width * height > 0x7fffffffLL) return false;
    cam.fx = intrinsic4[0]; cam.fy = intrinsic4[1]; cam.cx = intrinsic4[2]; cam.cy = intrinsic4[3];
    cam.depth_min = depth_range3[0]; cam.depth_max = depth_range3[1]; cam.accuracy = depth_range3[2];
    cam.width = width; cam.height = height;
    for (int k = 0; k < 8; ++k)
        for (int j = 0; j < 3; ++j) cam.corner[k][j] = corner_points[k * 3 + j];
    return true;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_frustum_planes(int views, const float *camera_to_world, const float *intrinsic4, const float *depth_range3,
                                  int width, int height, const float *corner_points, float *corners, float *normals,
                                  pn2_stream_t stream)
{
    ProjCamera cam;
    if (views < 0 || !fill_camera(cam, intrinsic4, depth_range3, width, height, corner_points)) return PN2_ERR_INVALID_ARGUMENT;
    if (views == 0) return PN2_OK;
    if (!camera_to_world) return PN2_ERR_INVALID_ARGUMENT;
    frustum_planes_kernel<<<ceil_div(views, 32), 32, 0, as_stream(stream)>>>(views, camera_to_world, cam, corners, normals);
    PN2_LAUNCH_CHECK("frustum_planes");
    return PN2_OK;
}

extern "C" int pn2_points_in_frustum(int n, const float *points, const float *corners, const float *normals,
                                     unsigned char *mask, int *count, pn2_stream_t stream)
{
    if (n < 0 || !count) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), as_stream(stream)));
    if (n == 0) return PN2_OK;
    if (!points || !corners || !normals) return PN2_ERR_INVALID_ARGUMENT;
    const int want = ceil_div(n, kProjThreads), cap = stream_sm_count(as_stream(stream)) * 8;
    points_in_frustum_kernel<<<want < cap ? want : cap, kProjThreads, 0, as_stream(stream)>>>(n, points, corners, normals, mask, count);
    PN2_LAUNCH_CHECK("points_in_frustum");
    return PN2_OK;
}

extern "C" size_t pn2_compute_projection_workspace_bytes(int views, int n)
{
    if (views < 0 || n < 0) return 0;
    const size_t nseg = (size_t)ceil_div(n > 0 ? n : 1, kProjSegment);
    return ((size_t)views * n + (size_t)views * nseg) * sizeof(int) + 256;
}

extern "C" int pn2_compute_projection(int views, int n, const float *points, const float *depth, const float *camera_to_world,
                                      const float *world_to_camera, const float *intrinsic4, const float *depth_range3,
                                      int width, int height, const float *corner_points, long long *indices_3d,
                                      long long *indices_2d, int *counts, void *workspace, size_t workspace_bytes,
                                      pn2_stream_t stream)
{
    ProjCamera cam;
    if (views < 0 || n < 0 || !fill_camera(cam, intrinsic4, depth_range3, width, height, corner_points)) return PN2_ERR_INVALID_ARGUMENT;
    if (views == 0) return PN2_OK;
    if (views > 65535 || !indices_3d || !indices_2d) return PN2_ERR_INVALID_ARGUMENT;
    if (n == 0) {          // the lists are just their zero counter
        PN2_CUDA_TRY(cudaMemsetAsync(indices_3d, 0, sizeof(long long) * views, as_stream(stream)));
        PN2_CUDA_TRY(cudaMemsetAsync(indices_2d, 0, sizeof(long long) * views, as_stream(stream)));
        if (counts) PN2_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * views, as_stream(stream)));
        return PN2_OK;
    }
    if (!points || !depth || !camera_to_world || !world_to_camera) return PN2_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < pn2_compute_projection_workspace_bytes(views, n)) return PN2_ERR_WORKSPACE;
    const int nseg = ceil_div(n, kProjSegment);
    int *pixmap = static_cast<int *>(workspace);
    int *segcount = pixmap + (size_t)views * n;
    const dim3 grid(nseg, views);
    projection_flags_kernel<<<grid, kProjThreads, 0, as_stream(stream)>>>(n, nseg, points, depth, camera_to_world, world_to_camera,
                                                                          cam, pixmap, segcount);
    PN2_LAUNCH_CHECK("projection_flags");
    projection_compact_kernel<<<grid, kProjThreads, 0, as_stream(stream)>>>(n, nseg, pixmap, segcount, indices_3d, indices_2d, counts);
    PN2_LAUNCH_CHECK("projection_compact");
    return PN2_OK;
}

// Across-view pooling of the back-projected features: out[pt, ch] = max(0, max over the views v that see point pt of
// label[v, ch, pixmap[v, pt]]) -- the running element-wise max over frames that produces the 128-d multiview input
// ("enet_feats_maxpool", lib/config.py:36; per frame: ProjectionHelper.project, lib/projection.py:257-279, into a
// zero-initialised per-point array).  One thread per point and CPB channels in registers; the (views, c, n) tensor that
// project() materialises per frame (1.6 GB for 64 views x 128 channels x 50 000 points) is never written.
template <int CPB>
__global__ void __launch_bounds__(kProjThreads)
project_maxpool_kernel(int views, int n, int c, int hw, const int *__restrict__ pixmap, const float *__restrict__ label,
                       float *__restrict__ out, int rows_layout)
{
    const int c0 = blockIdx.y * CPB;
    const int pt = blockIdx.x * kProjThreads + threadIdx.x;
    if (pt >= n) return;
    const int nc = min(CPB, c - c0);
    float acc[CPB];
#pragma unroll
    for (int j = 0; j < CPB; ++j) acc[j] = 0.f;
    for (int v = 0; v < views; ++v) {
        const int pix = pixmap[(size_t)v * n + pt];
        if (pix < 0) continue;
        const float *lab = label + ((size_t)v * c + c0) * hw + pix;
#pragma unroll
        for (int j = 0; j < CPB; ++j)
            if (j < nc) acc[j] = fmaxf(acc[j], __ldg(lab + (size_t)j * hw));
    }
    if (rows_layout) {
        float *o = out + (size_t)pt * c + c0;
#pragma unroll
        for (int j = 0; j < CPB; ++j)
            if (j < nc) o[j] = acc[j];
    } else {
#pragma unroll
        for (int j = 0; j < CPB; ++j)
            if (j < nc) out[(size_t)(c0 + j) * n + pt] = acc[j];
    }
}

extern "C" size_t pn2_project_workspace_bytes(int views, int n)
{
    if (views < 0 || n < 0) return 0;
    return (size_t)views * n * sizeof(int) + 256;
}

extern "C" int pn2_project(int views, int c, int hw, int n, const float *label, const long long *indices_3d,
                           const long long *indices_2d, float *out, int *status, void *workspace, size_t workspace_bytes,
                           pn2_stream_t stream)
{
    if (views < 0 || c < 1 || hw < 1 || n < 0 || !status) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    if (views == 0 || n == 0) return PN2_OK;
    if (views > 65535 || !label || !indices_3d || !indices_2d || !out) return PN2_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < pn2_project_workspace_bytes(views, n)) return PN2_ERR_WORKSPACE;
    int *pixmap = static_cast<int *>(workspace);
    PN2_CUDA_TRY(cudaMemsetAsync(pixmap, 0xff, sizeof(int) * (size_t)views * n, as_stream(stream)));
    const int blocks = ceil_div(n, kProjThreads);
    project_map_kernel<<<dim3(blocks < 64 ? blocks : 64, views), kProjThreads, 0, as_stream(stream)>>>(n, hw, indices_3d, indices_2d,
                                                                                                       pixmap, status);
    PN2_LAUNCH_CHECK("project_map");
    constexpr int kCpb = 32;
    if (ceil_div(c, kCpb) > 65535) return PN2_ERR_INVALID_ARGUMENT;
    // (a variant with four points per thread and 16-byte stores measured 0.70 ms against 0.38 ms for this one at
    //  64 views x 128 channels x 50 000 points: nearly every warp then holds a correspondence and runs the gather path)
    project_dense_kernel<kCpb><<<dim3(blocks, ceil_div(c, kCpb), views), kProjThreads, 0, as_stream(stream)>>>(n, c, hw, pixmap, label, out);
    PN2_LAUNCH_CHECK("project_dense");
    return PN2_OK;
}

extern "C" int pn2_project_maxpool(int views, int c, int hw, int n, const float *label, const long long *indices_3d,
                                   const long long *indices_2d, float *out, int rows_layout, int *status, void *workspace,
                                   size_t workspace_bytes, pn2_stream_t stream)
{
    if (views < 0 || c < 1 || hw < 1 || n < 0 || !status) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), as_stream(stream)));
    if (n == 0) return PN2_OK;
    if (views > 65535 || !out || (views > 0 && (!label || !indices_3d || !indices_2d))) return PN2_ERR_INVALID_ARGUMENT;
    if (views > 0 && (!workspace || workspace_bytes < pn2_project_workspace_bytes(views, n))) return PN2_ERR_WORKSPACE;
    int *pixmap = static_cast<int *>(workspace);
    const int blocks = ceil_div(n, kProjThreads);
    if (views > 0) {
        PN2_CUDA_TRY(cudaMemsetAsync(pixmap, 0xff, sizeof(int) * (size_t)views * n, as_stream(stream)));
        project_map_kernel<<<dim3(blocks < 64 ? blocks : 64, views), kProjThreads, 0, as_stream(stream)>>>(n, hw, indices_3d,
                                                                                                           indices_2d, pixmap, status);
        PN2_LAUNCH_CHECK("project_map");
    }
    constexpr int kCpb = 16;
    if (ceil_div(c, kCpb) > 65535) return PN2_ERR_INVALID_ARGUMENT;
    project_maxpool_kernel<kCpb><<<dim3(blocks, ceil_div(c, kCpb)), kProjThreads, 0, as_stream(stream)>>>(views, n, c, hw, pixmap, label,
                                                                                                        out, rows_layout ? 1 : 0);
    PN2_LAUNCH_CHECK("project_maxpool");
    return PN2_OK;
}
