// train_rows.cu -- training-mode BatchNorm + ReLU (+ max-pool over nsample) on channel-last rows, forward and
// backward, for the backbone's training step (BASELINE.json config 4).
//
// Reference: SharedMLP = Conv2d(1x1, no bias) -> BatchNorm2d -> ReLU per layer (lib/pointnet2/pytorch_utils.py:11-36,
// 67-121), then F.max_pool2d over nsample (lib/pointnet2/pointnet2_modules.py:259-262).  BatchNorm2d in training mode
// normalises each channel over B x npoint x nsample with the biased variance and eps inside the square root; on the
// row layout (row = one sample of one centre, columns = channels) that is a per-column statistic over all rows.
//
// All of it is HBM-bound element-wise work over matrices of 2^17 .. 2^20 rows x 64 .. 256 channels (SA1: 268 MB per
// activation at B = 8), so the only thing that matters is how many times each matrix crosses HBM:
//
//              PyTorch ops (stats, transform, relu | relu', bn reduce, bn elemt | amax, eq, mul)     here
//   forward    3 reads + 2 writes per layer (+ 1 read, 1 small write for the pool)          2 reads + 1 write
//              layer followed by the pool: the activation is never written                  2 reads
//   backward   6 reads + 2 writes per layer (+ ~4 passes for the pool's mask/scale)         4 reads + 1 write
//              layer followed by the pool: gradient is non-zero at one row per (centre,     1 read + 1 write
//              channel), so the two reductions touch only those rows
//
// The ReLU mask is recomputed from x (pre > 0 <=> x*a + b > 0): nothing but x is kept for the backward pass.
// Statistics are accumulated per thread in fp32 over <= a few hundred rows, merged in double in shared memory and
// written as one partial per CTA (no atomics: deterministic); a small finalize kernel adds the partials in a fixed
// order and forms mean / variance / running statistics (forward) or the parameter gradients and the coefficients of the
// input gradient (backward) in double -- two launches per layer and direction, no host synchronisation.
//
// Layout contract: x, y, dy, dx are (rows, c) fp32 row-major, c % 4 == 0, 16-byte aligned; a, b, k1..k3 are (c,) fp32.
#include "common.cuh"

namespace pn2 {

constexpr int kTrThreads = 256;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Merges per-thread fp32 partials (two float4 per thread: s = sums, q = second sums) over the row lanes of a CTA in
// double and writes them to this CTA's slot of `partials` ((gridDim.x, 2c) doubles: no atomics, no memset; the
// finalize kernels add the slots up in a fixed order, so the statistics are run-to-run deterministic).
// Threads are laid out as tid = rlane * (c/4) + cq.
__device__ __forceinline__ void merge_partials(float4 s, float4 q, int c, double *__restrict__ partials)
{
    __shared__ double red[2][kTrThreads * 4];   // [which][thread x 4 channels]  (16 KB)
    const int tid = threadIdx.x, tpr = c >> 2, cq = tid % tpr;
    const int nlanes = kTrThreads / tpr;
    double *r0 = &red[0][tid * 4], *r1 = &red[1][tid * 4];
    r0[0] = s.x; r0[1] = s.y; r0[2] = s.z; r0[3] = s.w;
    r1[0] = q.x; r1[1] = q.y; r1[2] = q.z; r1[3] = q.w;
    __syncthreads();
    // the first row lane's threads sum their column over the lanes
    if (tid < tpr) {
        double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
        for (int l = 0; l < nlanes; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0[j] += red[0][(l * tpr + cq) * 4 + j];
                a1[j] += red[1][(l * tpr + cq) * 4 + j];
            }
        double *out = partials + (size_t)blockIdx.x * 2 * c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[cq * 4 + j] = a0[j];
            out[c + cq * 4 + j] = a1[j];
        }
    }
}

// partials[cta] = [sum x | sum x^2] over this CTA's rows
__global__ void __launch_bounds__(kTrThreads)
bn_stats_kernel(long long rows, int c, const float *__restrict__ x, double *__restrict__ partials)
{
    const int tpr = c >> 2, cq = threadIdx.x % tpr, rlane = threadIdx.x / tpr, rpi = kTrThreads / tpr;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    const long long stride = (long long)gridDim.x * rpi;
    long long r = (long long)blockIdx.x * rpi + rlane;
    for (; r + 3 * stride < rows; r += 4 * stride) {          // four independent loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg4(x + (r + u * stride) * c + cq * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
            q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y);
            q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
        }
    }
    for (; r < rows; r += stride) {
        const float4 v = ldg4(x + r * c + cq * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    merge_partials(s, q, c, partials);
}

constexpr int kFinThreads = 1024;

// Sum of the partial slots for the 32 channels of this CTA (finalize kernels: one CTA per 32 channels, so that no single
// SM has to pull all (nparts x 2c) doubles through its L2 port -- one CTA for 256 channels took 15 us): thread
// (lane, j) adds the slots lane, lane + 32, ... of channel 32*blockIdx.x + j in a fixed order, lane 0 adds the 32
// results.  Returns true on the threads that hold a channel's totals (s1, s2).
constexpr int kFinChannels = 32;
__device__ __forceinline__ bool sum_partials(int c, int nparts, const double *__restrict__ partials, int &ch, double &s1,
                                             double &s2)
{
    __shared__ double fin[2][kFinThreads];
    constexpr int L = kFinThreads / kFinChannels;
    const int lane = threadIdx.x / kFinChannels, j = threadIdx.x % kFinChannels;
    ch = blockIdx.x * kFinChannels + j;
    double a1 = 0.0, a2 = 0.0;
    if (ch < c) {
        int p = lane;
        for (; p + 3 * L < nparts; p += 4 * L) {           // eight independent loads in flight (the slots sit in L2)
            double u[4], v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u[k] = partials[(size_t)(p + k * L) * 2 * c + ch];
                v[k] = partials[(size_t)(p + k * L) * 2 * c + c + ch];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { a1 += u[k]; a2 += v[k]; }
        }
        for (; p < nparts; p += L) {
            a1 += partials[(size_t)p * 2 * c + ch];
            a2 += partials[(size_t)p * 2 * c + c + ch];
        }
    }
    fin[0][threadIdx.x] = a1;
    fin[1][threadIdx.x] = a2;
    __syncthreads();
    if (lane != 0 || ch >= c) return false;
    s1 = 0.0; s2 = 0.0;
    for (int l = 0; l < L; ++l) { s1 += fin[0][l * kFinChannels + j]; s2 += fin[1][l * kFinChannels + j]; }
    return true;
}

// Forward finalize (one CTA per 32 channels): batch mean / biased variance from the partials, a = gamma/sqrt(var+eps), b = beta - mean*a,
// stat = [mean | 1/sqrt(var+eps)] in double for the backward pass, and nn.BatchNorm2d's running-statistics update
// (running_var takes the unbiased variance) when the buffers are given.
__global__ void __launch_bounds__(kFinThreads)
bn_finalize_kernel(int c, int nparts, const double *__restrict__ partials, double rows, double eps,
                   const float *__restrict__ weight, const float *__restrict__ bias, float momentum,
                   float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ a,
                   float *__restrict__ b, double *__restrict__ stat)
{
    int ch;
    double s1, s2;
    if (sum_partials(c, nparts, partials, ch, s1, s2)) {
        const double mean = s1 / rows;
        const double var = fmax(s2 / rows - mean * mean, 0.0);
        const double invstd = rsqrt(var + eps);
        const double g = (double)weight[ch];
        a[ch] = (float)(g * invstd);
        b[ch] = (float)((double)bias[ch] - mean * g * invstd);
        stat[ch] = mean;
        stat[c + ch] = invstd;
        if (running_mean) running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mean);
        if (running_var)
            running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * var * (rows / fmax(rows - 1.0, 1.0)));
    }
}

// Backward finalize (one CTA per 32 channels): from [sum g | sum g*x] the parameter gradients and the coefficients of
// dx = gamma*s*(g - mean(g) - x_hat*mean(g*x_hat)) = k1*g + k2 + k3*x   (s = 1/sqrt(var+eps), x_hat = (x - mean)*s)
__global__ void __launch_bounds__(kFinThreads)
bn_bwd_finalize_kernel(int c, int nparts, const double *__restrict__ partials, double rows, const double *__restrict__ stat,
                       const float *__restrict__ weight, float *__restrict__ k1, float *__restrict__ k2,
                       float *__restrict__ k3, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    int ch;
    double s1, s2;
    if (sum_partials(c, nparts, partials, ch, s1, s2)) {
        const double mean = stat[ch], invstd = stat[c + ch], g = (double)weight[ch];
        const double sum_gxhat = (s2 - mean * s1) * invstd;
        const double m1 = s1 / rows, m2 = sum_gxhat / rows, gs = g * invstd;
        k1[ch] = (float)gs;
        k2[ch] = (float)(-gs * m1 + gs * invstd * mean * m2);
        k3[ch] = (float)(-gs * invstd * m2);
        dgamma[ch] = (float)sum_gxhat;
        dbeta[ch] = (float)s1;
    }
}

// y = max(x * a + b, 0)
__global__ void __launch_bounds__(kTrThreads)
bn_relu_apply_kernel(long long n4, int c4, const float *__restrict__ x, const float *__restrict__ a,
                     const float *__restrict__ b, float *__restrict__ y)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const float4 v = ldg4(x + i * 4), av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
        st4(y + i * 4, make_float4(fmaxf(fmaf(v.x, av.x, bv.x), 0.f), fmaxf(fmaf(v.y, av.y, bv.y), 0.f),
                                   fmaxf(fmaf(v.z, av.z, bv.z), 0.f), fmaxf(fmaf(v.w, av.w, bv.w), 0.f)));
    }
}

// pooled[g, ch] = max_j max(x[g*ns + j, ch] * a + b, 0); arg[g, ch] = first j attaining the pre-activation maximum
// (max_pool2d keeps the first maximum; ReLU is monotone, so the argmax of the pre-activation is an argmax of the output)
__global__ void __launch_bounds__(kTrThreads)
bn_relu_pool_kernel(long long groups, int ns, int c, const float *__restrict__ x, const float *__restrict__ a,
                    const float *__restrict__ b, float *__restrict__ pooled, unsigned char *__restrict__ arg)
{
    const int tpr = c >> 2;
    const long long total = groups * tpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / tpr;
        const int cq = (int)(i - g * tpr);
        const float4 av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
        const float *p = x + (g * ns) * c + cq * 4;
        float m[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
        int am[4] = {0, 0, 0, 0};
        for (int j = 0; j < ns; ++j) {
            const float4 v = ldg4(p + (long long)j * c);
            const float pre[4] = {fmaf(v.x, av.x, bv.x), fmaf(v.y, av.y, bv.y), fmaf(v.z, av.z, bv.z), fmaf(v.w, av.w, bv.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (pre[k] > m[k]) { m[k] = pre[k]; am[k] = j; }      // strict: the first maximum stays (NaN never wins)
        }
        st4(pooled + g * c + cq * 4, make_float4(fmaxf(m[0], 0.f), fmaxf(m[1], 0.f), fmaxf(m[2], 0.f), fmaxf(m[3], 0.f)));
        *reinterpret_cast<uchar4 *>(arg + g * c + cq * 4) =
            make_uchar4((unsigned char)am[0], (unsigned char)am[1], (unsigned char)am[2], (unsigned char)am[3]);
    }
}

// g = dy where x*a + b > 0, else 0:  partials[cta] = [sum g | sum g*x] over this CTA's rows
__global__ void __launch_bounds__(kTrThreads)
bn_relu_bwd_reduce_kernel(long long rows, int c, const float *__restrict__ dy, const float *__restrict__ x,
                          const float *__restrict__ a, const float *__restrict__ b, double *__restrict__ partials)
{
    const int tpr = c >> 2, cq = threadIdx.x % tpr, rlane = threadIdx.x / tpr, rpi = kTrThreads / tpr;
    const float4 av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    auto acc = [&](const float4 &v, const float4 &d) {
        const float g0 = fmaf(v.x, av.x, bv.x) > 0.f ? d.x : 0.f, g1 = fmaf(v.y, av.y, bv.y) > 0.f ? d.y : 0.f;
        const float g2 = fmaf(v.z, av.z, bv.z) > 0.f ? d.z : 0.f, g3 = fmaf(v.w, av.w, bv.w) > 0.f ? d.w : 0.f;
        s.x += g0; s.y += g1; s.z += g2; s.w += g3;
        q.x = fmaf(g0, v.x, q.x); q.y = fmaf(g1, v.y, q.y); q.z = fmaf(g2, v.z, q.z); q.w = fmaf(g3, v.w, q.w);
    };
    const long long stride = (long long)gridDim.x * rpi;
    long long r = (long long)blockIdx.x * rpi + rlane;
    for (; r + stride < rows; r += 2 * stride) {              // two rows (four loads) in flight per thread
        const float4 v0 = ldg4(x + r * c + cq * 4), d0 = ldg4(dy + r * c + cq * 4);
        const float4 v1 = ldg4(x + (r + stride) * c + cq * 4), d1 = ldg4(dy + (r + stride) * c + cq * 4);
        acc(v0, d0);
        acc(v1, d1);
    }
    for (; r < rows; r += stride) acc(ldg4(x + r * c + cq * 4), ldg4(dy + r * c + cq * 4));
    merge_partials(s, q, c, partials);
}

// dx = k1 * g + k2 + k3 * x   (g as above)
__global__ void __launch_bounds__(kTrThreads)
bn_relu_bwd_apply_kernel(long long n4, int c4, const float *__restrict__ dy, const float *__restrict__ x,
                         const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ k1,
                         const float *__restrict__ k2, const float *__restrict__ k3, float *__restrict__ dx)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const float4 v = ldg4(x + i * 4), d = ldg4(dy + i * 4);
        const float4 av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
        const float4 u1 = ldg4(k1 + cq * 4), u2 = ldg4(k2 + cq * 4), u3 = ldg4(k3 + cq * 4);
        float4 o;
        o.x = fmaf(u3.x, v.x, u2.x) + (fmaf(v.x, av.x, bv.x) > 0.f ? u1.x * d.x : 0.f);
        o.y = fmaf(u3.y, v.y, u2.y) + (fmaf(v.y, av.y, bv.y) > 0.f ? u1.y * d.y : 0.f);
        o.z = fmaf(u3.z, v.z, u2.z) + (fmaf(v.z, av.z, bv.z) > 0.f ? u1.z * d.z : 0.f);
        o.w = fmaf(u3.w, v.w, u2.w) + (fmaf(v.w, av.w, bv.w) > 0.f ? u1.w * d.w : 0.f);
        st4(dx + i * 4, o);
    }
}

// pooled variant of the reduction: the gradient reaches one row per (group, channel)
__global__ void __launch_bounds__(kTrThreads)
bn_relu_pool_bwd_reduce_kernel(long long groups, int ns, int c, const float *__restrict__ dpooled,
                               const float *__restrict__ x, const unsigned char *__restrict__ arg,
                               const float *__restrict__ a, const float *__restrict__ b, double *__restrict__ partials)
{
    const int tpr = c >> 2, cq = threadIdx.x % tpr, rlane = threadIdx.x / tpr, rpi = kTrThreads / tpr;
    const float4 av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
    const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long g = (long long)blockIdx.x * rpi + rlane; g < groups; g += (long long)gridDim.x * rpi) {
        const float4 d4 = ldg4(dpooled + g * c + cq * 4);
        const uchar4 j4 = *reinterpret_cast<const uchar4 *>(arg + g * c + cq * 4);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const int j[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v = __ldg(x + (g * ns + j[k]) * c + cq * 4 + k);
            const float gk = fmaf(v, aa[k], bb[k]) > 0.f ? d[k] : 0.f;
            s[k] += gk;
            q[k] = fmaf(gk, v, q[k]);
        }
    }
    merge_partials(make_float4(s[0], s[1], s[2], s[3]), make_float4(q[0], q[1], q[2], q[3]), c, partials);
}

// dx[g*ns + j, ch] = k2 + k3 * x + (j == arg[g, ch] and pre > 0 ? k1 * dpooled[g, ch] : 0)
__global__ void __launch_bounds__(kTrThreads)
bn_relu_pool_bwd_apply_kernel(long long groups, int ns, int c, const float *__restrict__ dpooled,
                              const float *__restrict__ x, const unsigned char *__restrict__ arg,
                              const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ k1,
                              const float *__restrict__ k2, const float *__restrict__ k3, float *__restrict__ dx)
{
    const int tpr = c >> 2;
    const long long total = groups * tpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / tpr;
        const int cq = (int)(i - g * tpr);
        const float4 av = ldg4(a + cq * 4), bv = ldg4(b + cq * 4);
        const float4 u1 = ldg4(k1 + cq * 4), u2 = ldg4(k2 + cq * 4), u3 = ldg4(k3 + cq * 4);
        const float4 d = ldg4(dpooled + g * c + cq * 4);
        const uchar4 j4 = *reinterpret_cast<const uchar4 *>(arg + g * c + cq * 4);
        const float *p = x + (g * ns) * c + cq * 4;
        float *o = dx + (g * ns) * c + cq * 4;
        for (int j = 0; j < ns; ++j) {
            const float4 v = ldg4(p + (long long)j * c);
            float4 r;
            r.x = fmaf(u3.x, v.x, u2.x) + ((j == j4.x && fmaf(v.x, av.x, bv.x) > 0.f) ? u1.x * d.x : 0.f);
            r.y = fmaf(u3.y, v.y, u2.y) + ((j == j4.y && fmaf(v.y, av.y, bv.y) > 0.f) ? u1.y * d.y : 0.f);
            r.z = fmaf(u3.z, v.z, u2.z) + ((j == j4.z && fmaf(v.z, av.z, bv.z) > 0.f) ? u1.z * d.z : 0.f);
            r.w = fmaf(u3.w, v.w, u2.w) + ((j == j4.w && fmaf(v.w, av.w, bv.w) > 0.f) ? u1.w * d.w : 0.f);
            st4(o + (long long)j * c, r);
        }
    }
}


// Grouped input matrix of an SA module, built in one pass: out[r] = [ (xyz[idx[r]] - centre[r / ns]) / radius | rows[idx[r]] ]
// (QueryAndGroup: pointnet2_utils.py:348-359 -- grouped_xyz -= centre, /= radius, xyz channels first).
// idx holds scene-local indices; r runs over (scene, centre, sample).  One warp per output row, consecutive lanes on
// consecutive channels (coalesced reads of the source row and writes of the output row).
__global__ void __launch_bounds__(kTrThreads)
group_rows_kernel(long long total_rows, int n, int per_scene, int ns, int c, int ld, const int *__restrict__ idx,
                  const float *__restrict__ xyz, const float *__restrict__ centres, const float *__restrict__ rows,
                  float radius, int normalize, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int w = c + 3;
    for (long long r = (long long)blockIdx.x * (kTrThreads / 32) + (threadIdx.x >> 5); r < total_rows;
         r += (long long)gridDim.x * (kTrThreads / 32)) {
        const long long scene = r / per_scene;
        const long long src = scene * n + __ldg(idx + r);
        float *o = out + r * w;
        if (lane < 3) {
            float d = __fsub_rn(__ldg(xyz + src * 3 + lane), __ldg(centres + (r / ns) * 3 + lane));
            // torch divides a CUDA tensor by a Python scalar as a multiplication by the fp32 reciprocal (BinaryDivTrueKernel)
            if (normalize) d = __fmul_rn(d, __fdiv_rn(1.f, radius));
            o[lane] = d;
        }
        const float *p = rows + src * ld;
        for (int j = lane; j < c; j += 32) o[3 + j] = __ldg(p + j);
    }
}

static bool shape_ok(long long rows, int c) { return rows >= 0 && c >= 4 && c % 4 == 0 && c <= 1024 && kTrThreads % (c / 4) == 0; }

// CTAs of a reduction over `rows` (= slots of its partials buffer): at most two per SM, never more than there are
// row groups.  Depends on the shape and the device only, so the caller can size the buffer up front.
static int reduce_grid(long long rows, int c, cudaStream_t s)
{
    const int rpi = kTrThreads / (c / 4);
    const long long want = (rows + rpi - 1) / rpi;
    return (int)max(1LL, min(want, (long long)stream_sm_count(s) * 2));
}

static int flat_grid(long long items, cudaStream_t s)
{
    const long long want = (items + kTrThreads - 1) / kTrThreads;
    return (int)max(1LL, min(want, (long long)stream_sm_count(s) * 16));
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_rows_bn_supported(long long rows, int c, int nsample)
{
    return shape_ok(rows, c) && nsample >= 0 && nsample <= 256 ? 1 : 0;
}

extern "C" size_t pn2_rows_bn_partials_bytes(long long rows, int c)
{
    if (!shape_ok(rows, c)) return 0;
    return sizeof(double) * 2 * c * (size_t)(kNumSMs * 4);      // upper bound of reduce_grid() on any partition
}

// partials: pn2_rows_bn_partials_bytes() bytes; *nparts receives the number of slots written
extern "C" int pn2_rows_bn_stats(long long rows, int c, const float *x, double *partials, int *nparts, pn2_stream_t stream)
{
    if (!shape_ok(rows, c) || rows < 1 || !partials || !nparts || !x) return PN2_ERR_INVALID_ARGUMENT;
    cudaStream_t s = as_stream(stream);
    *nparts = reduce_grid(rows, c, s);
    bn_stats_kernel<<<*nparts, kTrThreads, 0, s>>>(rows, c, x, partials);
    PN2_LAUNCH_CHECK("rows_bn_stats");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_finalize(int c, int nparts, const double *partials, long long rows, double eps, const float *weight,
                                    const float *bias, float momentum, float *running_mean, float *running_var, float *a,
                                    float *b, double *stat, pn2_stream_t stream)
{
    if (c < 1 || nparts < 1 || rows < 1 || !partials || !weight || !bias || !a || !b || !stat) return PN2_ERR_INVALID_ARGUMENT;
    bn_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, kFinThreads, 0, as_stream(stream)>>>(c, nparts, partials, (double)rows, eps, weight, bias, momentum,
                                                          running_mean, running_var, a, b, stat);
    PN2_LAUNCH_CHECK("rows_bn_finalize");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_bwd_finalize(int c, int nparts, const double *partials, long long rows, const double *stat,
                                        const float *weight, float *k1, float *k2, float *k3, float *dgamma, float *dbeta,
                                        pn2_stream_t stream)
{
    if (c < 1 || nparts < 1 || rows < 1 || !partials || !stat || !weight || !k1 || !k2 || !k3 || !dgamma || !dbeta)
        return PN2_ERR_INVALID_ARGUMENT;
    bn_bwd_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, kFinThreads, 0, as_stream(stream)>>>(c, nparts, partials, (double)rows, stat, weight, k1, k2, k3,
                                                              dgamma, dbeta);
    PN2_LAUNCH_CHECK("rows_bn_bwd_finalize");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_apply(long long rows, int c, const float *x, const float *a, const float *b, float *y,
                                      pn2_stream_t stream)
{
    if (!shape_ok(rows, c) || !a || !b || (rows > 0 && (!x || !y))) return PN2_ERR_INVALID_ARGUMENT;
    if (rows == 0) return PN2_OK;
    cudaStream_t s = as_stream(stream);
    const long long n4 = rows * (c / 4);
    bn_relu_apply_kernel<<<flat_grid(n4, s), kTrThreads, 0, s>>>(n4, c / 4, x, a, b, y);
    PN2_LAUNCH_CHECK("rows_bn_relu_apply");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_pool(long long groups, int nsample, int c, const float *x, const float *a, const float *b,
                                     float *pooled, unsigned char *arg, pn2_stream_t stream)
{
    if (!shape_ok(groups, c) || nsample < 1 || nsample > 256 || !a || !b || (groups > 0 && (!x || !pooled || !arg)))
        return PN2_ERR_INVALID_ARGUMENT;
    if (groups == 0) return PN2_OK;
    cudaStream_t s = as_stream(stream);
    bn_relu_pool_kernel<<<flat_grid(groups * (c / 4), s), kTrThreads, 0, s>>>(groups, nsample, c, x, a, b, pooled, arg);
    PN2_LAUNCH_CHECK("rows_bn_relu_pool");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_bwd_reduce(long long rows, int c, const float *dy, const float *x, const float *a,
                                           const float *b, double *partials, int *nparts, pn2_stream_t stream)
{
    if (!shape_ok(rows, c) || rows < 1 || !partials || !nparts || !a || !b || !x || !dy) return PN2_ERR_INVALID_ARGUMENT;
    cudaStream_t s = as_stream(stream);
    *nparts = reduce_grid(rows, c, s);
    bn_relu_bwd_reduce_kernel<<<*nparts, kTrThreads, 0, s>>>(rows, c, dy, x, a, b, partials);
    PN2_LAUNCH_CHECK("rows_bn_relu_bwd_reduce");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_bwd_apply(long long rows, int c, const float *dy, const float *x, const float *a,
                                          const float *b, const float *k1, const float *k2, const float *k3, float *dx,
                                          pn2_stream_t stream)
{
    if (!shape_ok(rows, c) || !a || !b || !k1 || !k2 || !k3 || (rows > 0 && (!x || !dy || !dx)))
        return PN2_ERR_INVALID_ARGUMENT;
    if (rows == 0) return PN2_OK;
    cudaStream_t s = as_stream(stream);
    const long long n4 = rows * (c / 4);
    bn_relu_bwd_apply_kernel<<<flat_grid(n4, s), kTrThreads, 0, s>>>(n4, c / 4, dy, x, a, b, k1, k2, k3, dx);
    PN2_LAUNCH_CHECK("rows_bn_relu_bwd_apply");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_pool_bwd_reduce(long long groups, int nsample, int c, const float *dpooled, const float *x,
                                                const unsigned char *arg, const float *a, const float *b, double *partials,
                                                int *nparts, pn2_stream_t stream)
{
    if (!shape_ok(groups, c) || groups < 1 || nsample < 1 || nsample > 256 || !partials || !nparts || !a || !b || !x || !dpooled ||
        !arg)
        return PN2_ERR_INVALID_ARGUMENT;
    cudaStream_t s = as_stream(stream);
    *nparts = reduce_grid(groups, c, s);
    bn_relu_pool_bwd_reduce_kernel<<<*nparts, kTrThreads, 0, s>>>(groups, nsample, c, dpooled, x, arg, a, b, partials);
    PN2_LAUNCH_CHECK("rows_bn_relu_pool_bwd_reduce");
    return PN2_OK;
}

extern "C" int pn2_rows_bn_relu_pool_bwd_apply(long long groups, int nsample, int c, const float *dpooled, const float *x,
                                               const unsigned char *arg, const float *a, const float *b, const float *k1,
                                               const float *k2, const float *k3, float *dx, pn2_stream_t stream)
{
    if (!shape_ok(groups, c) || nsample < 1 || nsample > 256 || !a || !b || !k1 || !k2 || !k3 ||
        (groups > 0 && (!x || !dpooled || !arg || !dx)))
        return PN2_ERR_INVALID_ARGUMENT;
    if (groups == 0) return PN2_OK;
    cudaStream_t s = as_stream(stream);
    bn_relu_pool_bwd_apply_kernel<<<flat_grid(groups * (c / 4), s), kTrThreads, 0, s>>>(groups, nsample, c, dpooled, x, arg, a, b,
                                                                                      k1, k2, k3, dx);
    PN2_LAUNCH_CHECK("rows_bn_relu_pool_bwd_apply");
    return PN2_OK;
}

// out (b*npoint*nsample, c+3) = [normalised relative xyz | gathered feature rows]; rows: (b, n, ld) with c <= ld
extern "C" int pn2_group_rows(int b, int n, int npoint, int nsample, int c, int ld, const int *idx, const float *xyz,
                              const float *new_xyz, const float *rows, float radius, int normalize_xyz, float *out,
                              pn2_stream_t stream)
{
    if (b < 0 || n < 1 || npoint < 0 || nsample < 1 || c < 0 || ld < c) return PN2_ERR_INVALID_ARGUMENT;
    const long long total = (long long)b * npoint * nsample;
    if (total == 0) return PN2_OK;
    if (!idx || !xyz || !new_xyz || !out || (c > 0 && !rows)) return PN2_ERR_INVALID_ARGUMENT;
    cudaStream_t s = as_stream(stream);
    const long long want = (total + kTrThreads / 32 - 1) / (kTrThreads / 32);
    const int grid = (int)max(1LL, min(want, (long long)stream_sm_count(s) * 16));
    group_rows_kernel<<<grid, kTrThreads, 0, s>>>(total, n, npoint * nsample, nsample, c, ld, idx, xyz, new_xyz, rows, radius,
                                                  normalize_xyz, out);
    PN2_LAUNCH_CHECK("group_rows");
    return PN2_OK;
}
