// mlp_f32.cu -- full-precision (fp32 FFMA) fused set-abstraction and feature-propagation
// layers.  This is the precision="fp32" arm of the fused path (BASELINE.json config 1:
// rtol 1e-3 against the reference); the tensor-core arm lives in sa_tc.cu.
//
// One CTA works on a tile of 64 rows (row = one neighbour sample of one centre for SA, one
// unknown point for FP).  It gathers the rows into shared memory (SA: recentred/normalised
// xyz + neighbour features, pointnet2_utils.py:348-359; FP: three_interpolate + concat,
// pointnet2_modules.py:399-413), runs every SharedMLP layer (1x1 conv with BatchNorm folded
// + ReLU, pytorch_utils.py:11-36,67-121) out of shared memory with a 4x4 register tile per
// thread, and reduces the last layer with the max over nsample (pointnet2_modules.py:259-262)
// before anything is written to HBM.  The (B, C+3, npoint, nsample) tensor the reference
// materialises three times per layer never exists.
#include "common.cuh"
#include <math_constants.h>

namespace pn2 {

constexpr int kRows = 64;
constexpr int kMlpThreads = 256;
constexpr int kMaxLayers = 4;

struct MlpDesc {
    int nlayers;
    int dims[kMaxLayers + 1];      // dims[0] = input width, dims[l+1] = output width of layer l
    int kpad[kMaxLayers];          // input width rounded up to 4
    int cpad[kMaxLayers];          // output width rounded up to 64
    long long w_off[kMaxLayers];   // float offsets into the image: Wt[kpad][cpad]
    long long b_off[kMaxLayers];   // bias[cpad]
    long long total;               // floats
    int ld[2];                     // row pitch (floats) of the two ping-pong activation buffers
};

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
// pitch: multiple of 4 (float4 rows) and == 4 mod 8 (rows 4 apart land 16 banks apart)
static inline int pitch_for(int width) { int p = round_up(width, 4); return (p % 8 == 0) ? p + 4 : p; }

static bool make_desc(int nlayers, const int *dims, MlpDesc *d)
{
    if (nlayers < 1 || nlayers > kMaxLayers || !dims) return false;
    d->nlayers = nlayers;
    long long off = 0;
    int w[2] = {0, 0};
    for (int l = 0; l <= nlayers; ++l) {
        if (dims[l] < 1) return false;
        d->dims[l] = dims[l];
        w[l & 1] = max(w[l & 1], l < nlayers ? round_up(dims[l], 4) : dims[l]);
        if (l > 0) w[l & 1] = max(w[l & 1], round_up(dims[l], 64));   // a layer writes its padded width
    }
    for (int l = 0; l < nlayers; ++l) {
        d->kpad[l] = round_up(dims[l], 4);
        d->cpad[l] = round_up(dims[l + 1], 64);
        d->w_off[l] = off; off += (long long)d->kpad[l] * d->cpad[l];
        d->b_off[l] = off; off += d->cpad[l];
    }
    d->total = off;
    d->ld[0] = pitch_for(w[0]);
    d->ld[1] = pitch_for(w[1]);
    return true;
}

static size_t act_smem_bytes(const MlpDesc &d, int rows = kRows) { return sizeof(float) * rows * (size_t)(d.ld[0] + d.ld[1]); }

// image[w_off + k*cpad + c] = w[c*cin + k]   (transposed, zero padded);  image[b_off + c] = bias[c]
__global__ void pack_layer_kernel(int cin, int cout, int kpad, int cpad, const float *__restrict__ w,
                                  const float *__restrict__ bias, float *__restrict__ wt,
                                  float *__restrict__ bt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kpad * cpad) {
        const int k = i / cpad, c = i % cpad;
        wt[i] = (k < cin && c < cout) ? __ldg(w + (size_t)c * cin + k) : 0.f;
    }
    if (i < cpad) bt[i] = (i < cout && bias) ? __ldg(bias + i) : 0.f;
}

// One layer over an R-row tile: out[r][c] = act(bias[c] + sum_k in[r][k] * wt[k][c]).
// 256 threads = R/4 row groups (4 rows) x 1024/R channel groups (4 channels); channels are covered
// in passes of 4096/R.  R = 64: a warp touches two row groups (broadcast LDS.128) and 16 consecutive
// float4 of one weight row; R = 16 (few rows, e.g. the FP layers: more CTAs, each streaming the
// weights once): a warp shares its rows and reads 512 contiguous bytes of the weight row.
template <int R>
__device__ __forceinline__ void mlp_layer(const float *__restrict__ in, int ldin, int kpad,
                                          const float *__restrict__ wt, const float *__restrict__ bias,
                                          int cpad, float *__restrict__ out, int ldout, bool relu)
{
    constexpr int CG = 1024 / R;                  // channel groups
    const int tr = threadIdx.x / CG, tc = threadIdx.x % CG;
    const float *a0 = in + (tr * 4) * ldin;
    for (int cb = 0; cb < cpad; cb += CG * 4) {
        const int c0 = cb + tc * 4;
        if (c0 >= cpad) break;
        float acc[4][4];
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + c0));
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
        const float *wp = wt + c0;
        for (int k = 0; k < kpad; k += 4) {
            float4 a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(a0 + i * ldin + k);
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)(k + q) * cpad));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[i][0] = fmaf(av[q], w[q].x, acc[i][0]);
                    acc[i][1] = fmaf(av[q], w[q].y, acc[i][1]);
                    acc[i][2] = fmaf(av[q], w[q].z, acc[i][2]);
                    acc[i][3] = fmaf(av[q], w[q].w, acc[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(out + (tr * 4 + i) * ldout + c0) = o;
        }
    }
}

// runs all layers; returns the buffer index (0/1) holding the last layer's output
template <int R>
__device__ __forceinline__ int run_mlp(const MlpDesc &d, const float *__restrict__ image, float *buf0,
                                       float *buf1)
{
    for (int l = 0; l < d.nlayers; ++l) {
        __syncthreads();
        const bool odd = l & 1;
        mlp_layer<R>(odd ? buf1 : buf0, odd ? d.ld[1] : d.ld[0], d.kpad[l], image + d.w_off[l], image + d.b_off[l],
                  d.cpad[l], odd ? buf0 : buf1, odd ? d.ld[0] : d.ld[1], true);
    }
    __syncthreads();
    return d.nlayers & 1;
}

// ---- fused SA layer -------------------------------------------------------------------
// table: channel-last feature rows, row (b, p) at table + (b*n + p)*ld, c floats used.
// work item = G centres (G*nsample <= 64 rows) or, when nsample > 64, one centre processed
// in chunks of 64 rows with a running max.
__global__ void __launch_bounds__(kMlpThreads)
sa_forward_f32_kernel(MlpDesc d, int n, int npoint, int nsample, int c, int ld, float inv_radius,
                      int use_xyz, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                      const float *__restrict__ table, const int *__restrict__ idx,
                      const float *__restrict__ image, float *__restrict__ out,
                      float *__restrict__ out_rows, int groups_per_scene, int G)
{
    extern __shared__ __align__(16) float act[];
    float *buf0 = act, *buf1 = act + kRows * d.ld[0];
    __shared__ float red[kRows * 4];   // running max for the nsample > 64 case (cout <= 256)

    const int bi = blockIdx.x / groups_per_scene;
    const int grp = blockIdx.x % groups_per_scene;
    const int centre0 = grp * G;
    const int ncent = min(G, npoint - centre0);
    const int cout = d.dims[d.nlayers];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xoff = use_xyz ? 3 : 0;
    const int nchunks = nsample > kRows ? ceil_div(nsample, kRows) : 1;

    for (int ch = 0; ch < nchunks; ++ch) {
        const int s0 = ch * kRows;
        const int rows = nsample > kRows ? min(kRows, nsample - s0) : ncent * nsample;
        __syncthreads();
        // gather: warp per row, lanes along channels (coalesced row reads, conflict-free writes)
        for (int r = warp; r < kRows; r += kMlpThreads / 32) {
            float *dst = buf0 + r * d.ld[0];
            if (r < rows) {
                const int g = nsample > kRows ? 0 : r / nsample;
                const int s = nsample > kRows ? s0 + r : r - g * nsample;
                const int centre = centre0 + g;
                const int nb = __ldg(idx + ((size_t)bi * npoint + centre) * nsample + s);
                if (use_xyz && lane < 3) {
                    const float pc = __ldg(xyz + ((size_t)bi * n + nb) * 3 + lane);
                    const float cc = __ldg(new_xyz + ((size_t)bi * npoint + centre) * 3 + lane);
                    dst[lane] = __fmul_rn(__fsub_rn(pc, cc), inv_radius);   // pointnet2_utils.py:350-352
                }
                const float *src = table + ((size_t)bi * n + nb) * ld;
                for (int k = lane; k < c; k += 32) dst[xoff + k] = __ldg(src + k);
                for (int k = xoff + c + lane; k < d.kpad[0]; k += 32) dst[k] = 0.f;
            } else {
                for (int k = lane; k < d.kpad[0]; k += 32) dst[k] = 0.f;
            }
        }
        const int ob = run_mlp<kRows>(d, image, buf0, buf1);
        const float *res = ob ? buf1 : buf0;
        const int ldr = d.ld[ob];
        // max over the samples of each centre (pointnet2_modules.py:259-262)
        if (nsample > kRows) {
            for (int cc = threadIdx.x; cc < cout; cc += kMlpThreads) {
                float v = ch == 0 ? -CUDART_INF_F : red[cc];
                for (int r = 0; r < rows; ++r) v = fmaxf(v, res[r * ldr + cc]);
                red[cc] = v;
                if (ch == nchunks - 1) {
                    out[((size_t)bi * cout + cc) * npoint + centre0] = v;
                    if (out_rows) out_rows[((size_t)bi * npoint + centre0) * cout + cc] = v;
                }
            }
        } else {
            for (int t = threadIdx.x; t < ncent * cout; t += kMlpThreads) {
                const int g = t / cout, cc = t - g * cout;
                float v = -CUDART_INF_F;
                for (int s = 0; s < nsample; ++s) v = fmaxf(v, res[(g * nsample + s) * ldr + cc]);
                out[((size_t)bi * cout + cc) * npoint + centre0 + g] = v;
                if (out_rows) out_rows[((size_t)bi * npoint + centre0 + g) * cout + cc] = v;
            }
        }
    }
}

// ---- fused FP layer -------------------------------------------------------------------
// known_rows (b, m, c_known) and skip_rows (b, n, c_skip) are channel-last; dist2/idx come from
// pn2_three_nn.  Weights follow pointnet2_modules.py:399-402 (sqrt, 1/(d+1e-8), normalise) and
// the interpolation is the reference's FMUL,FFMA,FFMA (interpolate_gpu.cu:90-99).
template <int R>
__global__ void __launch_bounds__(kMlpThreads)
fp_forward_f32_kernel(MlpDesc d, int n, int m, int c_known, int c_skip, const float *__restrict__ dist2,
                      const int *__restrict__ idx, const float *__restrict__ known_rows,
                      const float *__restrict__ skip_rows, const float *__restrict__ image,
                      float *__restrict__ out, float *__restrict__ out_rows, int tiles_per_scene)
{
    extern __shared__ __align__(16) float act[];
    float *buf0 = act, *buf1 = act + R * d.ld[0];
    const int bi = blockIdx.x / tiles_per_scene;
    const int row0 = (blockIdx.x % tiles_per_scene) * R;
    const int rows = min(R, n - row0);
    const int cout = d.dims[d.nlayers];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int r = warp; r < R; r += kMlpThreads / 32) {
        float *dst = buf0 + r * d.ld[0];
        if (r < rows) {
            const size_t j = (size_t)bi * n + row0 + r;
            int kend = 0;
            if (c_known > 0) {
                const float d1 = __fsqrt_rn(__ldg(dist2 + j * 3)), d2 = __fsqrt_rn(__ldg(dist2 + j * 3 + 1)),
                            d3 = __fsqrt_rn(__ldg(dist2 + j * 3 + 2));
                const float r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)), r2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f)),
                            r3 = __fdiv_rn(1.0f, __fadd_rn(d3, 1e-8f));
                const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
                const float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
                const float *f1 = known_rows + ((size_t)bi * m + __ldg(idx + j * 3)) * c_known;
                const float *f2 = known_rows + ((size_t)bi * m + __ldg(idx + j * 3 + 1)) * c_known;
                const float *f3 = known_rows + ((size_t)bi * m + __ldg(idx + j * 3 + 2)) * c_known;
                for (int k = lane; k < c_known; k += 32)
                    dst[k] = interp3(__ldg(f1 + k), w1, __ldg(f2 + k), w2, __ldg(f3 + k), w3);
                kend = c_known;
            }
            if (c_skip > 0) {
                const float *sk = skip_rows + j * c_skip;
                for (int k = lane; k < c_skip; k += 32) dst[kend + k] = __ldg(sk + k);
                kend += c_skip;
            }
            for (int k = kend + lane; k < d.kpad[0]; k += 32) dst[k] = 0.f;
        } else {
            for (int k = lane; k < d.kpad[0]; k += 32) dst[k] = 0.f;
        }
    }
    const int ob = run_mlp<R>(d, image, buf0, buf1);
    const float *res = ob ? buf1 : buf0;
    const int ldr = d.ld[ob];
    // out (b, cout, n): rows fastest so that stores coalesce; out_rows (b, n, cout): channels fastest
    for (int t = threadIdx.x; t < cout * R; t += kMlpThreads) {
        const int cc = t / R, r = t % R;
        if (r < rows) out[((size_t)bi * cout + cc) * n + row0 + r] = res[r * ldr + cc];
    }
    if (out_rows)
        for (int t = threadIdx.x; t < rows * cout; t += kMlpThreads) {
            const int r = t / cout, cc = t - r * cout;
            out_rows[((size_t)bi * n + row0 + r) * cout + cc] = res[r * ldr + cc];
        }
}

// (b, c, n) -> (b, n, c) through a 32x32 shared-memory tile
__global__ void transpose_to_rows_kernel(int c, int n, const float *__restrict__ src, float *__restrict__ dst)
{
    __shared__ float t[32][33];
    const size_t bi = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int cc = c0 + i, nn = n0 + threadIdx.x;
        t[i][threadIdx.x] = (cc < c && nn < n) ? __ldg(src + (bi * c + cc) * n + nn) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int nn = n0 + i, cc = c0 + threadIdx.x;
        if (cc < c && nn < n) dst[(bi * n + nn) * c + cc] = t[threadIdx.x][i];
    }
}

static const int kMaxSmem = 220 * 1024;

}  // namespace pn2

using namespace pn2;

extern "C" size_t pn2_mlp_f32_image_bytes(int nlayers, const int *dims)
{
    MlpDesc d;
    if (!make_desc(nlayers, dims, &d)) return 0;
    return sizeof(float) * (size_t)d.total;
}

extern "C" int pn2_mlp_f32_supported(int nlayers, const int *dims)
{
    MlpDesc d;
    if (!make_desc(nlayers, dims, &d)) return 0;
    return act_smem_bytes(d) <= (size_t)kMaxSmem ? 1 : 0;
}

extern "C" int pn2_mlp_f32_pack(int nlayers, const int *dims, const float *const *w, const float *const *bias,
                                void *image, pn2_stream_t stream)
{
    MlpDesc d;
    if (!make_desc(nlayers, dims, &d) || !w || !image) return PN2_ERR_INVALID_ARGUMENT;
    float *img = static_cast<float *>(image);
    for (int l = 0; l < nlayers; ++l) {
        if (!w[l]) return PN2_ERR_INVALID_ARGUMENT;
        const int total = d.kpad[l] * d.cpad[l];
        pack_layer_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
            d.dims[l], d.dims[l + 1], d.kpad[l], d.cpad[l], w[l], bias ? bias[l] : nullptr, img + d.w_off[l],
            img + d.b_off[l]);
        PN2_LAUNCH_CHECK("mlp_f32_pack");
    }
    return PN2_OK;
}

extern "C" int pn2_rows_from_channels(int b, int c, int n, const float *src, float *dst, pn2_stream_t stream)
{
    if (b < 0 || c < 0 || n < 0 || b > 65535) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || c == 0 || n == 0) return PN2_OK;
    if (!src || !dst) return PN2_ERR_INVALID_ARGUMENT;
    dim3 grid(ceil_div(n, 32), ceil_div(c, 32), b), block(32, 8);
    if (grid.y > 65535) return PN2_ERR_INVALID_ARGUMENT;
    transpose_to_rows_kernel<<<grid, block, 0, as_stream(stream)>>>(c, n, src, dst);
    PN2_LAUNCH_CHECK("rows_from_channels");
    return PN2_OK;
}

extern "C" int pn2_sa_forward_f32(int b, int n, int npoint, int nsample, int c, const float *table, int ld,
                                  int use_xyz, float inv_radius, const float *xyz, const float *new_xyz,
                                  const int *idx, int nlayers, const int *dims, const void *image,
                                  float *out, float *out_rows, pn2_stream_t stream)
{
    MlpDesc d;
    if (!make_desc(nlayers, dims, &d)) return PN2_ERR_INVALID_ARGUMENT;
    if (b < 0 || n < 1 || npoint < 0 || nsample < 1 || c < 0 || ld < c) return PN2_ERR_INVALID_ARGUMENT;
    if (d.dims[0] != c + (use_xyz ? 3 : 0)) return PN2_ERR_INVALID_ARGUMENT;
    if (nsample > kRows && d.dims[nlayers] > kRows * 4) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || npoint == 0) return PN2_OK;
    if (!xyz || !new_xyz || !idx || !image || !out || (c > 0 && !table)) return PN2_ERR_INVALID_ARGUMENT;
    const size_t smem = act_smem_bytes(d);
    if (smem > (size_t)kMaxSmem) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaFuncSetAttribute(sa_forward_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    const int G = nsample > kRows ? 1 : kRows / nsample;
    const int groups = ceil_div(npoint, G);
    sa_forward_f32_kernel<<<(unsigned)((long long)b * groups), kMlpThreads, smem, as_stream(stream)>>>(
        d, n, npoint, nsample, c, ld, inv_radius, use_xyz, xyz, new_xyz, table, idx,
        static_cast<const float *>(image), out, out_rows, groups, G);
    PN2_LAUNCH_CHECK("sa_forward_f32");
    return PN2_OK;
}

extern "C" int pn2_fp_forward_f32(int b, int n, int m, int c_known, int c_skip, const float *dist2,
                                  const int *idx, const float *known_rows, const float *skip_rows,
                                  int nlayers, const int *dims, const void *image, float *out,
                                  float *out_rows, pn2_stream_t stream)
{
    MlpDesc d;
    if (!make_desc(nlayers, dims, &d)) return PN2_ERR_INVALID_ARGUMENT;
    if (b < 0 || n < 0 || m < 0 || c_known < 0 || c_skip < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (d.dims[0] != c_known + c_skip) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || n == 0) return PN2_OK;
    if (!image || !out || (c_known > 0 && (!dist2 || !idx || !known_rows || m < 1)) || (c_skip > 0 && !skip_rows))
        return PN2_ERR_INVALID_ARGUMENT;
    if (act_smem_bytes(d) > (size_t)kMaxSmem) return PN2_ERR_INVALID_ARGUMENT;
    // few rows (the backbone's FP layers have 512 / 1024 points per scene): 16-row tiles give 4x the CTAs
    const bool small = (long long)b * ceil_div(n, kRows) < 4 * kNumSMs;
    if (small) {
        const size_t smem = act_smem_bytes(d, 16);
        PN2_CUDA_TRY(cudaFuncSetAttribute(fp_forward_f32_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        const int tiles = ceil_div(n, 16);
        fp_forward_f32_kernel<16><<<(unsigned)((long long)b * tiles), kMlpThreads, smem, as_stream(stream)>>>(
            d, n, m, c_known, c_skip, dist2, idx, known_rows, skip_rows, static_cast<const float *>(image), out,
            out_rows, tiles);
    } else {
        const size_t smem = act_smem_bytes(d);
        PN2_CUDA_TRY(cudaFuncSetAttribute(fp_forward_f32_kernel<kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
        const int tiles = ceil_div(n, kRows);
        fp_forward_f32_kernel<kRows><<<(unsigned)((long long)b * tiles), kMlpThreads, smem, as_stream(stream)>>>(
            d, n, m, c_known, c_skip, dist2, idx, known_rows, skip_rows, static_cast<const float *>(image), out,
            out_rows, tiles);
    }
    PN2_LAUNCH_CHECK("fp_forward_f32");
    return PN2_OK;
}
