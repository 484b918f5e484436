// ops_basic.cu -- gather / group / three_interpolate and their gradients.
//
// These are pure HBM-bound index-driven copies.  The reference runs each with
// ONE block per batch element (grid = B; sampling_gpu.cu:25, group_points_gpu.cu:35,
// interpolate_gpu.cu:107), i.e. on B of 148 SMs.  Here every op is laid out so
// that consecutive threads own consecutive elements of the contiguous output
// dimension (coalesced stores), the index is read once and reused over a
// channel chunk, and the grid covers (elements x channel-chunks x batch).
#include "common.cuh"

namespace pn2 {

constexpr int kThreads = 256;
constexpr int kChunk = 8;   // channels per thread: one idx load feeds 8 gathers

// out[b,c,j] = points[b,c,idx[b,j]]            (reference: sampling_gpu.cu:8-20)
__global__ void __launch_bounds__(kThreads)
gather_points_kernel(int c, int n, int m, const float *__restrict__ points,
                     const int *__restrict__ idx, float *__restrict__ out)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    const size_t bi = blockIdx.z;
    const int a = __ldg(idx + bi * m + j);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l)
        out[(bi * c + l) * m + j] = __ldg(points + (bi * c + l) * n + a);
}

// grad_points[b,c,idx[b,j]] += grad_out[b,c,j]  (reference: sampling_gpu.cu:34-47)
__global__ void __launch_bounds__(kThreads)
gather_points_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                          const int *__restrict__ idx, float *__restrict__ grad_points)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    const size_t bi = blockIdx.z;
    const int a = __ldg(idx + bi * m + j);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l)
        atomicAdd(grad_points + (bi * c + l) * n + a, __ldg(grad_out + (bi * c + l) * m + j));
}

// out[b,c,t] = points[b,c,idx[b,t]], t = j*nsample + k   (reference: group_points_gpu.cu:8-28)
__global__ void __launch_bounds__(kThreads)
group_points_kernel(int c, int n, long long total, const float *__restrict__ points,
                    const int *__restrict__ idx, float *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const size_t bi = blockIdx.z;
    const int a = __ldg(idx + bi * total + t);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l)
        out[(bi * c + l) * total + t] = __ldg(points + (bi * c + l) * n + a);
}

// grad_points[b,c,idx[b,t]] += grad_out[b,c,t]   (reference: group_points_gpu.cu:43-64)
__global__ void __launch_bounds__(kThreads)
group_points_grad_kernel(int c, int n, long long total, const float *__restrict__ grad_out,
                         const int *__restrict__ idx, float *__restrict__ grad_points)
{
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const size_t bi = blockIdx.z;
    const int a = __ldg(idx + bi * total + t);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l)
        atomicAdd(grad_points + (bi * c + l) * n + a, __ldg(grad_out + (bi * c + l) * total + t));
}

// out[b,l,j] = p[l,i1]*w1 + p[l,i2]*w2 + p[l,i3]*w3 in the reference build's contraction order
// (reference: interpolate_gpu.cu:72-101)
__global__ void __launch_bounds__(kThreads)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= n) return;
    const size_t bi = blockIdx.z;
    const size_t o3 = (bi * n + j) * 3;
    const int i1 = __ldg(idx + o3), i2 = __ldg(idx + o3 + 1), i3 = __ldg(idx + o3 + 2);
    const float w1 = __ldg(weight + o3), w2 = __ldg(weight + o3 + 1), w3 = __ldg(weight + o3 + 2);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l) {
        const float *p = points + (bi * c + l) * m;
        out[(bi * c + l) * n + j] = interp3(__ldg(p + i1), w1, __ldg(p + i2), w2, __ldg(p + i3), w3);
    }
}

// grad_points[b,l,i_t] += grad_out[b,l,j] * w_t   (reference: interpolate_gpu.cu:116-143)
__global__ void __launch_bounds__(kThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= n) return;
    const size_t bi = blockIdx.z;
    const size_t o3 = (bi * n + j) * 3;
    const int i1 = __ldg(idx + o3), i2 = __ldg(idx + o3 + 1), i3 = __ldg(idx + o3 + 2);
    const float w1 = __ldg(weight + o3), w2 = __ldg(weight + o3 + 1), w3 = __ldg(weight + o3 + 2);
    const int c0 = blockIdx.y * kChunk;
    const int c1 = min(c, c0 + kChunk);
    for (int l = c0; l < c1; ++l) {
        const float g = __ldg(grad_out + (bi * c + l) * n + j);
        float *gp = grad_points + (bi * c + l) * m;
        atomicAdd(gp + i1, __fmul_rn(g, w1));
        atomicAdd(gp + i2, __fmul_rn(g, w2));
        atomicAdd(gp + i3, __fmul_rn(g, w3));
    }
}

static bool bad_dims(int b, int c, long long inner)
{
    return b < 0 || c < 0 || inner < 0 || b > 65535 || ceil_div(c, kChunk) > 65535;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                                 float *out, pn2_stream_t stream)
{
    if (bad_dims(b, c, m) || n < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || m == 0) return PN2_OK;
    if (!points || !idx || !out) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(m, kThreads), ceil_div(c, kChunk), b);
    gather_points_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, n, m, points, idx, out);
    PN2_LAUNCH_CHECK("gather_points");
    return PN2_OK;
}

extern "C" int pn2_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                      const int *idx, float *grad_points, pn2_stream_t stream)
{
    if (bad_dims(b, c, m) || n < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || m == 0) return PN2_OK;
    if (!grad_out || !idx || !grad_points) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(m, kThreads), ceil_div(c, kChunk), b);
    gather_points_grad_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, n, m, grad_out, idx,
                                                                        grad_points);
    PN2_LAUNCH_CHECK("gather_points_grad");
    return PN2_OK;
}

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                const int *idx, float *out, pn2_stream_t stream)
{
    const long long total = (long long)npoints * nsample;
    if (bad_dims(b, c, total) || n < 0 || npoints < 0 || nsample < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || total == 0) return PN2_OK;
    if (!points || !idx || !out) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(total, kThreads), ceil_div(c, kChunk), b);
    group_points_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, n, total, points, idx, out);
    PN2_LAUNCH_CHECK("group_points");
    return PN2_OK;
}

extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int *idx, float *grad_points,
                                     pn2_stream_t stream)
{
    const long long total = (long long)npoints * nsample;
    if (bad_dims(b, c, total) || n < 0 || npoints < 0 || nsample < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || total == 0) return PN2_OK;
    if (!grad_out || !idx || !grad_points) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(total, kThreads), ceil_div(c, kChunk), b);
    group_points_grad_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, n, total, grad_out, idx,
                                                                       grad_points);
    PN2_LAUNCH_CHECK("group_points_grad");
    return PN2_OK;
}

extern "C" int pn2_three_interpolate(int b, int c, int m, int n, const float *points,
                                     const int *idx, const float *weight, float *out,
                                     pn2_stream_t stream)
{
    if (bad_dims(b, c, n) || m < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || n == 0) return PN2_OK;
    if (!points || !idx || !weight || !out) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(n, kThreads), ceil_div(c, kChunk), b);
    three_interpolate_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, m, n, points, idx, weight,
                                                                       out);
    PN2_LAUNCH_CHECK("three_interpolate");
    return PN2_OK;
}

extern "C" int pn2_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int *idx, const float *weight, float *grad_points,
                                          pn2_stream_t stream)
{
    if (bad_dims(b, c, n) || m < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || n == 0) return PN2_OK;
    if (!grad_out || !idx || !weight || !grad_points) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(n, kThreads), ceil_div(c, kChunk), b);
    three_interpolate_grad_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(c, n, m, grad_out, idx,
                                                                            weight, grad_points);
    PN2_LAUNCH_CHECK("three_interpolate_grad");
    return PN2_OK;
}
