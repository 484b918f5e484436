// head_tc.cu -- the linear + GELU that consumes the visual tokens.
//
// Reference: SIG3D.scene_feat_linear = Sequential(Linear(256, 768), GELU())  (situation3d/models/sqa_module.py:180-183),
// applied to the (B, 256, 256) tokens right after the positional embedding was added (:344).  In the reference this is a
// cuBLAS fp32 GEMM followed by an element-wise GELU kernel over the (B, 256, 768) result; here it is one tcgen05 kernel
// (bf16 operands, fp32 accumulation, bias + exact-erf GELU in the epilogue, fp32 output):
//   warps 0-3    convert the CTA's 128 token rows fp32 -> bf16 into the K-major swizzled A operand (once), then stream
//                the 128-row weight tiles of this CTA's output columns through a two-stage ring with cp.async
//   warp 4       one tcgen05.mma chain per weight tile, accumulators double buffered in TMEM
//   warps 8-15   epilogue (the erf is the expensive part): tcgen05.ld -> + bias -> GELU -> 256 bytes of an output row
//                per thread per tile
// grid = (row tiles, column splits): the A operand is built once per CTA and reused for all of its weight tiles.
#include "tc_common.cuh"

namespace pn2 {

constexpr int kHeadTile = 128;       // rows per CTA (MMA M) and output columns per weight tile (MMA N)
constexpr int kHeadLoaders = 128;     // warps 0-3
constexpr int kHeadWarpMma = 4;
constexpr int kHeadEpiWarp0 = 8;      // warps 8-15: TMEM lane quarter = warp % 4, column half = (warp - 8) / 4
constexpr int kHeadThreads = 512;

struct HeadParams {
    long long rows;
    int k, n, ntiles_n, tiles_per_cta;
    const float *x;
    const unsigned char *image;      // ntiles_n tiles of [128 rows][k] bf16, K-major swizzled
    const float *bias;
    float *out;
};

__global__ void head_pack_weights_kernel(int k, int n, const float *__restrict__ w, unsigned char *__restrict__ image)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * k) return;
    const int j = i / k, e = i - j * k;
    const size_t tile_bytes = kop_bytes(kHeadTile, k);
    *reinterpret_cast<__nv_bfloat16 *>(image + (size_t)(j / kHeadTile) * tile_bytes + kop_chunk_off(kHeadTile, k, j % kHeadTile, e >> 3) +
                                       (e & 7) * 2) = __float2bfloat16_rn(w[i]);
}

__device__ __forceinline__ float gelu_erf_f(float v) { return v * 0.5f * (1.0f + erff(v * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(kHeadThreads, 1)
head_tc_kernel(const HeadParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int K = p.k;
    const uint32_t t_bytes = (kop_bytes(kHeadTile, K) + 1023u) & ~1023u;
    unsigned char *as = base, *wb = base + t_bytes;                       // weight ring: wb, wb + t_bytes
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wb + 2 * t_bytes);      // a_full, w_full[2], w_empty[2], d_full[2], d_empty[2]
    const uint32_t bar = smem_u32(mbar);
    const uint32_t b_a_full = bar, b_w_full = bar + 8, b_w_empty = bar + 24, b_d_full = bar + 40, b_d_empty = bar + 56;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 9);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc_mbar_init(b_a_full, kHeadLoaders / 32);
        for (int i = 0; i < 2; ++i) {
            tc_mbar_init(b_w_full + 8u * i, kHeadLoaders / 32);
            tc_mbar_init(b_w_empty + 8u * i, 1);
            tc_mbar_init(b_d_full + 8u * i, 1);
            tc_mbar_init(b_d_empty + 8u * i, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<2 * kHeadTile>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const long long row0 = (long long)blockIdx.x * kHeadTile;
    const int j0 = blockIdx.y * p.tiles_per_cta, nj = min(p.tiles_per_cta, p.ntiles_n - j0);

    if (warp < kHeadLoaders / 32) {
        // A operand: 128 rows x K floats, 16-byte loads, converted in registers
        const int per_row = K / 4, total = kHeadTile * per_row;
        for (int f = tid; f < total; f += kHeadLoaders) {
            const int r = f / per_row, c = f - r * per_row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < p.rows) v = __ldg(reinterpret_cast<const float4 *>(p.x + (size_t)(row0 + r) * K + c * 4));
            *reinterpret_cast<uint2 *>(as + kop_chunk_off(kHeadTile, K, r, c >> 1) + (c & 1) * 8) =
                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_a_full) : "memory");
        const uint32_t chunks = kop_bytes(kHeadTile, K) / 16;
        for (int j = 0; j < nj; ++j) {
            const int st = j & 1, u = j >> 1;
            if (u > 0) tc_mbar_wait(b_w_empty + 8u * st, (u - 1) & 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(p.image + (size_t)(j0 + j) * kop_bytes(kHeadTile, K));
            const uint32_t dst = smem_u32(wb + st * t_bytes);
            for (uint32_t i = tid; i < chunks; i += kHeadLoaders) cp_async16(dst + i * 16, src + i);
            cp_async_wait_all();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_w_full + 8u * st) : "memory");
        }
    } else if (warp == kHeadWarpMma) {
        const uint32_t idesc = umma_idesc(kHeadTile, kHeadTile);
        const uint32_t elected = elect_one();
        tc_mbar_wait(b_a_full, 0);
        for (int j = 0; j < nj; ++j) {
            const int st = j & 1, u = j >> 1;
            tc_mbar_wait(b_w_full + 8u * st, u & 1);
            if (u > 0) tc_mbar_wait(b_d_empty + 8u * st, (u - 1) & 1);
            fence_proxy_async();
            tc_fence_after();
            issue_gemm(tmem + st * kHeadTile, smem_u32(as), kHeadTile, 0, smem_u32(wb + st * t_bytes), kHeadTile, 0, K, idesc,
                       elected);
            if (elected) {
                umma_commit(b_d_full + 8u * st);
                umma_commit(b_w_empty + 8u * st);
            }
            __syncwarp();
        }
    } else if (warp >= kHeadEpiWarp0) {
        const int quarter = warp & 3, half = (warp - kHeadEpiWarp0) >> 2, row = quarter * 32 + lane;
        const uint32_t my_tmem = tmem + ((uint32_t)(quarter * 32) << 16);
        const long long grow = row0 + row;
        for (int j = 0; j < nj; ++j) {
            const int st = j & 1, u = j >> 1;
            tc_mbar_wait(b_d_full + 8u * st, u & 1);
            tc_fence_after();
            const int col0 = (j0 + j) * kHeadTile;
#pragma unroll
            for (int c0 = half * (kHeadTile / 2); c0 < (half + 1) * (kHeadTile / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(my_tmem + st * kHeadTile + c0, v);
                if (grow < p.rows) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out + (size_t)grow * p.n + col0 + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + c0 + q * 4));
                        dst[q] = make_float4(gelu_erf_f(__uint_as_float(v[4 * q]) + b.x), gelu_erf_f(__uint_as_float(v[4 * q + 1]) + b.y),
                                             gelu_erf_f(__uint_as_float(v[4 * q + 2]) + b.z), gelu_erf_f(__uint_as_float(v[4 * q + 3]) + b.w));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_d_empty + 8u * st) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<2 * kHeadTile>(tmem);
}

static uint32_t head_smem_bytes(int k) { return 1024u + 3u * ((kop_bytes(kHeadTile, k) + 1023u) & ~1023u) + 9u * 8u + 16u; }

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_linear_gelu_tc_supported(int k, int n)
{
    if (k < 64 || k % 64 || n < kHeadTile || n % kHeadTile) return 0;
    return head_smem_bytes(k) <= 227u * 1024u ? 1 : 0;
}

extern "C" size_t pn2_linear_gelu_tc_weight_image_bytes(int k, int n)
{
    return pn2_linear_gelu_tc_supported(k, n) ? (size_t)(n / kHeadTile) * kop_bytes(kHeadTile, k) : 0;
}

extern "C" int pn2_linear_gelu_tc_pack_weights(int k, int n, const float *w, void *image, pn2_stream_t stream)
{
    if (!pn2_linear_gelu_tc_supported(k, n) || !w || !image) return PN2_ERR_INVALID_ARGUMENT;
    head_pack_weights_kernel<<<ceil_div((long long)n * k, 256), 256, 0, as_stream(stream)>>>(k, n, w, static_cast<unsigned char *>(image));
    PN2_LAUNCH_CHECK("linear_gelu_tc_pack_weights");
    return PN2_OK;
}

/* out (rows, n) f32 = GELU(x (rows, k) f32 . W^T + bias), W given as the packed image */
extern "C" int pn2_linear_gelu_tc_forward(long long rows, int k, int n, const float *x, const void *weight_image,
                                          const float *bias, float *out, pn2_stream_t stream)
{
    if (rows < 0 || !pn2_linear_gelu_tc_supported(k, n)) return PN2_ERR_INVALID_ARGUMENT;
    if (rows == 0) return PN2_OK;
    if (!x || !weight_image || !bias || !out) return PN2_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(bias) % 16)
        return PN2_ERR_INVALID_ARGUMENT;
    HeadParams p;
    p.rows = rows; p.k = k; p.n = n; p.ntiles_n = n / kHeadTile;
    p.x = x; p.image = static_cast<const unsigned char *>(weight_image); p.bias = bias; p.out = out;
    const int mtiles = (int)((rows + kHeadTile - 1) / kHeadTile);
    // split the output columns over CTAs until the grid fills the SMs (each CTA rebuilds the A operand)
    const int sms = stream_sm_count(as_stream(stream));
    int split = 1;
    while (split < p.ntiles_n && mtiles * split * 2 <= sms && p.ntiles_n % (split * 2) == 0) split *= 2;
    p.tiles_per_cta = p.ntiles_n / split;
    const uint32_t smem = head_smem_bytes(k);
    PN2_CUDA_TRY(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_tc_kernel<<<dim3(mtiles, split), kHeadThreads, smem, as_stream(stream)>>>(p);
    PN2_LAUNCH_CHECK("linear_gelu_tc_forward");
    return PN2_OK;
}
