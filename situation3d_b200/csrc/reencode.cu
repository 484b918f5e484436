// reencode.cu -- situation-conditioned re-encoding of visual tokens.
//
// Reference semantics:
//   situation3d/utils/temp.py:42-80,86-97   7-D situation (t, quat xyzw) -> 4x4, p' = R p + t
//   situation3d/models/sqa_module.py:274-278 pos_embed = Linear(2,h) -> GELU -> Linear(h,d)
//   situation3d/models/sqa_module.py:319-321 tokens + pos_embed(positions_xy)
//   situation3d/models/sqa_module.py:328-336 Gaussian location prior over the tokens
// The reference spreads this over ~25 tiny PyTorch kernels; here it is one launch for the
// embedding (+ one warp-sized launch for the normalised prior): a CTA takes 32 tokens,
// transforms their positions, builds the h hidden activations in shared memory and streams
// W2 through shared memory in 32-row slabs so that both operands of the d x h product are
// read conflict-free; tokens are read once and the sum is written once.
#include "common.cuh"

namespace pn2 {

constexpr int kTok = 32;
constexpr int kReThreads = 256;
constexpr int kSlab = 32;

__device__ __forceinline__ float gelu_erf(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

// rotation of temp.py:56-70 (x2 - y2 - z2 + w2 diagonal; no normalisation of the quaternion)
__device__ __forceinline__ void quat_matrix(const float *s, float R[9], float t[3])
{
    t[0] = s[0]; t[1] = s[1]; t[2] = s[2];
    const float x = s[3], y = s[4], z = s[5], w = s[6];
    const float x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    const float xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = x2 - y2 - z2 + w2; R[1] = 2.f * (xy - zw);      R[2] = 2.f * (xz + yw);
    R[3] = 2.f * (xy + zw);      R[4] = -x2 + y2 - z2 + w2; R[5] = 2.f * (yz - xw);
    R[6] = 2.f * (xz - yw);      R[7] = 2.f * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

__global__ void __launch_bounds__(kReThreads)
reencode_kernel(int t, int d, int h, int mode, const float *__restrict__ tokens,
                const float *__restrict__ positions, const float *__restrict__ situation,
                const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                const float *__restrict__ b2, float *__restrict__ out, float *__restrict__ new_pos,
                int groups_per_scene)
{
    extern __shared__ __align__(16) float re_smem[];
    // hid[kTok][hp] | slab[kSlab][dpitch] | pxy[kTok][2]
    const int hp = (h + 3) / 4 * 4;
    float *hid = re_smem;
    float *slab = hid + kTok * hp;
    float *pxy = slab + kSlab * (kReThreads + 1);

    const int bi = blockIdx.x / groups_per_scene;
    const int tok0 = (blockIdx.x % groups_per_scene) * kTok;
    const int ntok = min(kTok, t - tok0);

    if (threadIdx.x < kTok) {
        const int i = threadIdx.x;
        float q[3] = {0.f, 0.f, 0.f};
        if (i < ntok) {
            const float *p = positions + ((size_t)bi * t + tok0 + i) * 3;
            float R[9], tr[3];
            quat_matrix(situation + (size_t)bi * 7, R, tr);
            const float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
            if (mode == 0) {          // p' = R p + t            (temp.py:90-97)
                q[0] = fmaf(pz, R[2], fmaf(py, R[1], px * R[0])) + tr[0];
                q[1] = fmaf(pz, R[5], fmaf(py, R[4], px * R[3])) + tr[1];
                q[2] = fmaf(pz, R[8], fmaf(py, R[7], px * R[6])) + tr[2];
            } else {                  // agent frame: p' = R^T (p - t)
                const float ux = px - tr[0], uy = py - tr[1], uz = pz - tr[2];
                q[0] = fmaf(uz, R[6], fmaf(uy, R[3], ux * R[0]));
                q[1] = fmaf(uz, R[7], fmaf(uy, R[4], ux * R[1]));
                q[2] = fmaf(uz, R[8], fmaf(uy, R[5], ux * R[2]));
            }
            if (new_pos) {
                float *np = new_pos + ((size_t)bi * t + tok0 + i) * 3;
                np[0] = q[0]; np[1] = q[1]; np[2] = q[2];
            }
        }
        pxy[2 * i] = q[0]; pxy[2 * i + 1] = q[1];
    }
    __syncthreads();
    // hidden layer: Linear(2, h) + exact GELU (sqa_module.py:274-276)
    for (int e = threadIdx.x; e < kTok * hp; e += kReThreads) {
        const int i = e / hp, j = e - i * hp;
        float v = 0.f;
        if (j < h) v = gelu_erf(fmaf(pxy[2 * i + 1], __ldg(w1 + 2 * j + 1), fmaf(pxy[2 * i], __ldg(w1 + 2 * j), __ldg(b1 + j))));
        hid[e] = v;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d0 = 0; d0 < d; d0 += kReThreads) {
        const int dd = d0 + threadIdx.x;
        float acc[kTok];
        const float bias = dd < d ? __ldg(b2 + dd) : 0.f;
#pragma unroll
        for (int i = 0; i < kTok; ++i) acc[i] = bias;
        for (int h0 = 0; h0 < hp; h0 += kSlab) {
            __syncthreads();
            // slab[j][c] = w2[d0 + c][h0 + j]: warp reads 32 consecutive h of one row (coalesced)
            for (int c = warp; c < kReThreads; c += kReThreads / 32) {
                const int row = d0 + c, col = h0 + lane;
                slab[lane * (kReThreads + 1) + c] = (row < d && col < h) ? __ldg(w2 + (size_t)row * h + col) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kSlab; j += 4) {
                if (h0 + j >= hp) break;
                const float wa = slab[(j + 0) * (kReThreads + 1) + threadIdx.x];
                const float wb = slab[(j + 1) * (kReThreads + 1) + threadIdx.x];
                const float wc = slab[(j + 2) * (kReThreads + 1) + threadIdx.x];
                const float wd = slab[(j + 3) * (kReThreads + 1) + threadIdx.x];
#pragma unroll
                for (int i = 0; i < kTok; ++i) {
                    const float4 hv = *reinterpret_cast<const float4 *>(hid + i * hp + h0 + j);
                    acc[i] = fmaf(hv.w, wd, fmaf(hv.z, wc, fmaf(hv.y, wb, fmaf(hv.x, wa, acc[i]))));
                }
            }
        }
        if (dd < d) {
#pragma unroll
            for (int i = 0; i < kTok; ++i)
                if (i < ntok) {
                    const size_t o = ((size_t)bi * t + tok0 + i) * d + dd;
                    out[o] = __ldg(tokens + o) + acc[i];
                }
        }
    }
}

// ---- v2: persistent CTAs, W2 resident in shared memory, 8 tokens x 8 outputs per thread on packed FFMA -------------
// The kernel above re-stages W2 (d x h fp32 = 128 KB for the reference's 256 x 128) for every 32 tokens and spends one
// shared-memory read per two FMAs.  Here a CTA loads W2 transposed ONCE (W2T[h][d]: a lane's outputs are two float4
// groups, conflict-free) and walks 64-token tiles: positions -> hidden activations (exact-erf GELU) -> the d x h
// product with each thread holding 8 tokens x 8 outputs in registers as fp32x2 pairs: per hidden unit 4 LDS.128 feed
// 32 fma.rn.f32x2 (64 FMAs).  tokens are read once, the sum is written once.  Used when W2T + the tile fit in 200 KB.
constexpr int kTok2 = 64;

__device__ __forceinline__ unsigned long long re_pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long re_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int D>
__global__ void __launch_bounds__(kReThreads, 1)
reencode_v2_kernel(int b, int t, int h, int mode, const float *__restrict__ tokens, const float *__restrict__ positions,
                   const float *__restrict__ situation, const float *__restrict__ w1, const float *__restrict__ b1,
                   const float *__restrict__ w2, const float *__restrict__ b2, float *__restrict__ out,
                   float *__restrict__ new_pos, int tiles_per_scene)
{
    static_assert(D == 256, "a lane owns outputs 4*lane..+3 and 128 + 4*lane..+3");
    extern __shared__ __align__(16) float re2_smem[];
    float *w2t = re2_smem;                       // [h][D]
    float *hid = w2t + (size_t)h * D;            // [h][kTok2]
    float *pxy = hid + (size_t)h * kTok2;        // [kTok2][2]
    float *w1s = pxy + 2 * kTok2;                // [h][2] then b1 [h]: the hidden layer's parameters, staged once
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 3 * h; i += kReThreads) w1s[i] = i < 2 * h ? __ldg(w1 + i) : __ldg(b1 + (i - 2 * h));
    // W2 (D, h) row-major -> W2T[h][D].  A warp takes 32 consecutive outputs dd and one 16-byte piece of their rows:
    // the shared-memory stores are conflict-free (consecutive dd), the global reads use half of every 32-byte sector
    // (out of L2, once per CTA); eight independent loads are in flight per thread.
    if ((h & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0) {
        const int pieces = h / 4, total = (D / 32) * pieces;          // (dd block, piece) items per warp-iteration
        for (int it0 = warp; it0 < total; it0 += 8 * (kReThreads / 32)) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * (kReThreads / 32);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < total) {
                    const int dblk = it / pieces, pc = it - dblk * pieces;
                    v[u] = __ldg(reinterpret_cast<const float4 *>(w2 + (size_t)(dblk * 32 + lane) * h) + pc);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * (kReThreads / 32);
                if (it < total) {
                    const int dblk = it / pieces, pc = it - dblk * pieces;
                    float *dst = w2t + (size_t)(4 * pc) * D + dblk * 32 + lane;
                    dst[0] = v[u].x; dst[D] = v[u].y; dst[2 * D] = v[u].z; dst[3 * D] = v[u].w;
                }
            }
        }
    } else {
        for (int i = tid; i < h * D; i += kReThreads) {
            const int dd = i % D, hh = i / D;
            w2t[hh * D + dd] = __ldg(w2 + (size_t)dd * h + hh);
        }
    }
    const float4 bias_lo = __ldg(reinterpret_cast<const float4 *>(b2) + lane);
    const float4 bias_hi = __ldg(reinterpret_cast<const float4 *>(b2 + 128) + lane);
    const int ntiles = b * tiles_per_scene;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int bi = tile / tiles_per_scene, tok0 = (tile - bi * tiles_per_scene) * kTok2;
        const int ntok = min(kTok2, t - tok0);
        __syncthreads();                         // W2T staged / previous tile's hid consumed
        if (tid < kTok2) {
            float q[3] = {0.f, 0.f, 0.f};
            if (tid < ntok) {
                const float *pp = positions + ((size_t)bi * t + tok0 + tid) * 3;
                float R[9], tr[3];
                quat_matrix(situation + (size_t)bi * 7, R, tr);
                const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
                if (mode == 0) {          // p' = R p + t            (temp.py:90-97)
                    q[0] = fmaf(pz, R[2], fmaf(py, R[1], px * R[0])) + tr[0];
                    q[1] = fmaf(pz, R[5], fmaf(py, R[4], px * R[3])) + tr[1];
                    q[2] = fmaf(pz, R[8], fmaf(py, R[7], px * R[6])) + tr[2];
                } else {                  // agent frame: p' = R^T (p - t)
                    const float ux = px - tr[0], uy = py - tr[1], uz = pz - tr[2];
                    q[0] = fmaf(uz, R[6], fmaf(uy, R[3], ux * R[0]));
                    q[1] = fmaf(uz, R[7], fmaf(uy, R[4], ux * R[1]));
                    q[2] = fmaf(uz, R[8], fmaf(uy, R[5], ux * R[2]));
                }
                if (new_pos) {
                    float *np = new_pos + ((size_t)bi * t + tok0 + tid) * 3;
                    np[0] = q[0]; np[1] = q[1]; np[2] = q[2];
                }
            }
            pxy[2 * tid] = q[0]; pxy[2 * tid + 1] = q[1];
        }
        __syncthreads();
        // hidden layer: Linear(2, h) + exact GELU (sqa_module.py:274-276), stored [h][token]
        for (int e = tid; e < h * kTok2; e += kReThreads) {
            const int i = e % kTok2, jj = e / kTok2;
            hid[e] = gelu_erf(fmaf(pxy[2 * i + 1], w1s[2 * jj + 1], fmaf(pxy[2 * i], w1s[2 * jj], w1s[2 * h + jj])));
        }
        __syncthreads();
        // product: warp -> tokens 8*warp .. 8*warp+7, lane -> outputs 4*lane..+3 and 128 + 4*lane..+3
        unsigned long long acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i][0] = re_pack2(bias_lo.x, bias_lo.y); acc[i][1] = re_pack2(bias_lo.z, bias_lo.w);
            acc[i][2] = re_pack2(bias_hi.x, bias_hi.y); acc[i][3] = re_pack2(bias_hi.z, bias_hi.w);
        }
#pragma unroll 2
        for (int jj = 0; jj < h; ++jj) {
            const float4 wl = *reinterpret_cast<const float4 *>(w2t + (size_t)jj * D + 4 * lane);
            const float4 wh = *reinterpret_cast<const float4 *>(w2t + (size_t)jj * D + 128 + 4 * lane);
            const float4 ha = *reinterpret_cast<const float4 *>(hid + (size_t)jj * kTok2 + 8 * warp);
            const float4 hb = *reinterpret_cast<const float4 *>(hid + (size_t)jj * kTok2 + 8 * warp + 4);
            const unsigned long long w0 = re_pack2(wl.x, wl.y), w1p = re_pack2(wl.z, wl.w), w2p = re_pack2(wh.x, wh.y),
                                     w3 = re_pack2(wh.z, wh.w);
            const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned long long hh2 = re_pack2(hv[i], hv[i]);
                acc[i][0] = re_fma2(hh2, w0, acc[i][0]); acc[i][1] = re_fma2(hh2, w1p, acc[i][1]);
                acc[i][2] = re_fma2(hh2, w2p, acc[i][2]); acc[i][3] = re_fma2(hh2, w3, acc[i][3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tk = 8 * warp + i;
            if (tk < ntok) {
                const size_t o = ((size_t)bi * t + tok0 + tk) * D;
                const float4 tl = __ldg(reinterpret_cast<const float4 *>(tokens + o) + lane);
                const float4 th = __ldg(reinterpret_cast<const float4 *>(tokens + o + 128) + lane);
                float a[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * q]), "=f"(a[2 * q + 1]) : "l"(acc[i][q]));
                reinterpret_cast<float4 *>(out + o)[lane] = make_float4(tl.x + a[0], tl.y + a[1], tl.z + a[2], tl.w + a[3]);
                reinterpret_cast<float4 *>(out + o + 128)[lane] = make_float4(th.x + a[4], th.y + a[5], th.z + a[6], th.w + a[7]);
            }
        }
    }
}

// prior[b, i] = w_i / sum_i w_i,  w_i = exp(-|p_xy - t_xy|^2 / (2 sigma^2))   (sqa_module.py:332-336)
__global__ void __launch_bounds__(256)
prior_kernel(int t, float sigma, const float *__restrict__ positions, const float *__restrict__ situation,
             float *__restrict__ prior)
{
    __shared__ float part[8];
    const size_t bi = blockIdx.x;
    const float tx = __ldg(situation + bi * 7), ty = __ldg(situation + bi * 7 + 1);
    const float denom = 2.f * sigma * sigma;
    float sum = 0.f;
    for (int i = threadIdx.x; i < t; i += blockDim.x) {
        const float *p = positions + (bi * t + i) * 3;
        const float dx = __ldg(p) - tx, dy = __ldg(p + 1) - ty;
        const float dist = sqrtf(fmaf(dy, dy, dx * dx));      // torch.norm(dim=2), then squared again
        const float w = expf(-(dist * dist) / denom);
        prior[bi * t + i] = w;
        sum += w;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    float total = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += part[w];
    for (int i = threadIdx.x; i < t; i += blockDim.x) prior[bi * t + i] = prior[bi * t + i] / total;
}

// kind 0: quaternions (b,4) xyzw -> (b,3,3), the "1 - 2(y^2+z^2)" form of sqa_module.py:12-30
// kind 1: rotation vectors (b,3) -> (b,3,3), Rodrigues with identity below 1e-6 (sqa_module.py:33-64)
// kind 2: situations (b,7) -> (b,4,4) = [R|t; 0 0 0 1], temp.py:42-80
__global__ void rotation_matrices_kernel(int b, int kind, const float *__restrict__ in, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    if (kind == 0) {
        const float x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2], w = in[4 * i + 3];
        float *R = out + 9 * (size_t)i;
        R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w);       R[2] = 2.f * (x * z + y * w);
        R[3] = 2.f * (x * y + z * w);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
        R[6] = 2.f * (x * z - y * w);       R[7] = 2.f * (y * z + x * w);       R[8] = 1.f - 2.f * (x * x + y * y);
    } else if (kind == 1) {
        const float vx = in[3 * i], vy = in[3 * i + 1], vz = in[3 * i + 2];
        const float theta = sqrtf(vx * vx + vy * vy + vz * vz);
        float *R = out + 9 * (size_t)i;
        if (theta < 1e-6f) {
            R[0] = 1.f; R[1] = 0.f; R[2] = 0.f; R[3] = 0.f; R[4] = 1.f; R[5] = 0.f; R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
            return;
        }
        const float ux = vx / theta, uy = vy / theta, uz = vz / theta;
        const float K[9] = {0.f, -uz, uy, uz, 0.f, -ux, -uy, ux, 0.f};
        const float s = sinf(theta), c1 = 1.f - cosf(theta);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const float kk = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
                R[3 * r + c] = (r == c ? 1.f : 0.f) + s * K[3 * r + c] + c1 * kk;
            }
    } else {
        float R[9], t[3];
        quat_matrix(in + 7 * (size_t)i, R, t);
        float *M = out + 16 * (size_t)i;
        for (int r = 0; r < 3; ++r) {
            M[4 * r] = R[3 * r]; M[4 * r + 1] = R[3 * r + 1]; M[4 * r + 2] = R[3 * r + 2]; M[4 * r + 3] = t[r];
        }
        M[12] = 0.f; M[13] = 0.f; M[14] = 0.f; M[15] = 1.f;
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_reencode_forward(int b, int t, int d, int h, int mode, float sigma, const float *tokens,
                                    const float *positions, const float *situation, const float *w1,
                                    const float *b1, const float *w2, const float *b2, float *out,
                                    float *new_pos, float *prior, pn2_stream_t stream)
{
    if (b < 0 || t < 0 || d < 1 || h < 1 || (mode != 0 && mode != 1)) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || t == 0) return PN2_OK;
    if (!tokens || !positions || !situation || !w1 || !b1 || !w2 || !b2 || !out) return PN2_ERR_INVALID_ARGUMENT;
    const size_t smem2 = sizeof(float) * ((size_t)h * 256 + (size_t)h * kTok2 + 2 * kTok2 + 3 * (size_t)h);
    const bool aligned = ((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0;
    if (d == 256 && smem2 <= 200 * 1024 && aligned) {
        // the reference's shape (Linear(128, 256)): W2 resident per CTA, persistent over 64-token tiles
        PN2_CUDA_TRY(cudaFuncSetAttribute(reencode_v2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int tiles_per_scene = ceil_div(t, kTok2);
        const long long ntiles = (long long)b * tiles_per_scene;
        const int grid = (int)(ntiles < stream_sm_count(as_stream(stream)) ? ntiles : stream_sm_count(as_stream(stream)));
        reencode_v2_kernel<256><<<grid, kReThreads, smem2, as_stream(stream)>>>(b, t, h, mode, tokens, positions, situation,
                                                                              w1, b1, w2, b2, out, new_pos, tiles_per_scene);
        PN2_LAUNCH_CHECK("reencode_v2");
    } else {
        const int hp = (h + 3) / 4 * 4;
        const size_t smem = sizeof(float) * ((size_t)kTok * hp + (size_t)kSlab * (kReThreads + 1) + 2 * kTok);
        if (smem > 200 * 1024) return PN2_ERR_INVALID_ARGUMENT;
        PN2_CUDA_TRY(cudaFuncSetAttribute(reencode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int groups = ceil_div(t, kTok);
        reencode_kernel<<<(unsigned)((long long)b * groups), kReThreads, smem, as_stream(stream)>>>(
            t, d, h, mode, tokens, positions, situation, w1, b1, w2, b2, out, new_pos, groups);
        PN2_LAUNCH_CHECK("reencode");
    }
    if (prior) {
        if (!(sigma > 0.f)) return PN2_ERR_INVALID_ARGUMENT;
        prior_kernel<<<b, 256, 0, as_stream(stream)>>>(t, sigma, positions, situation, prior);
        PN2_LAUNCH_CHECK("reencode_prior");
    }
    return PN2_OK;
}

static int rotation_matrices(int b, int kind, const float *in, float *out, pn2_stream_t stream)
{
    if (b < 0) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0) return PN2_OK;
    if (!in || !out) return PN2_ERR_INVALID_ARGUMENT;
    rotation_matrices_kernel<<<ceil_div(b, 128), 128, 0, as_stream(stream)>>>(b, kind, in, out);
    PN2_LAUNCH_CHECK("rotation_matrices");
    return PN2_OK;
}

extern "C" int pn2_quaternions_to_rotation_matrices(int b, const float *quats, float *out, pn2_stream_t stream)
{
    return rotation_matrices(b, 0, quats, out, stream);
}

extern "C" int pn2_rotation_vectors_to_matrices(int b, const float *rotvecs, float *out, pn2_stream_t stream)
{
    return rotation_matrices(b, 1, rotvecs, out, stream);
}

extern "C" int pn2_situation_matrices(int b, const float *situation, float *out, pn2_stream_t stream)
{
    return rotation_matrices(b, 2, situation, out, stream);
}
