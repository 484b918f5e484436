// fp_tc.cu -- fused feature-propagation layer on the tensor cores (tcgen05 / TMEM).
//
// Replaces PointnetFPModule.forward (pointnet2_modules.py:399-421) in eval mode after three_nn:
//   sqrt -> 1/(d + 1e-8) -> normalise            pointnet2_modules.py:399-402
//   three_interpolate                            interpolate_gpu.cu:72-101
//   cat([interpolated, skip], dim=1)             pointnet2_modules.py:412-416
//   SharedMLP (2 x [1x1 conv, BatchNorm, ReLU])  pytorch_utils.py:11-36,67-121
// with ONE kernel.  A CTA owns a tile of 128 unknown points (= MMA M).  The concatenated input row
// (c_known interpolated + c_skip skip channels, 512 for the backbone) never exists in HBM: it is produced
// 64 channels at a time straight into the K-major, 128B-swizzled A-operand layout -- the interpolated
// part by the row's thread (three 128-byte row reads, fp32 FMUL/FFMA/FFMA in the reference's order, bf16
// round), the skip part by cp.async -- while the matching 64-column slab of W1 (32 KB) streams in next to
// it.  A two-stage ring lets the production of slab k+1 overlap the tcgen05.mma of slab k; accumulators
// (128 x 256 fp32) live in TMEM.  Layer 2 reuses the ring for the W2 slabs and reads the bf16 activations
// of layer 1 from shared memory.  The epilogue writes the reference's (B, C, n) fp32 tensor (coalesced over
// points) and the bf16 channel-last rows the next layer gathers from.
#include "tc_common.cuh"

namespace pn2 {

constexpr int kFpTile = 128;
constexpr int kFpThreads = 128;
constexpr int kSlab = 64;                                  // K columns per ring stage

struct FpTcShape {
    int c_known, c_skip, c1, c2, k0;
    uint32_t w1_bytes, w2_bytes;                          // image sections; then (c1 + c2) fp32 biases
    uint32_t stage_a_bytes, stage_w_bytes, a1_bytes, smem_bytes;
};

static bool make_fp_shape(int c_known, int c_skip, int c1, int c2, FpTcShape *s)
{
    if (c_known < kSlab || c_known % kSlab || c_skip < 0 || c_skip % kSlab) return false;
    if (c1 < 16 || c1 > 256 || c1 % kSlab || c2 < 16 || c2 > 256 || c2 % 16) return false;
    s->c_known = c_known; s->c_skip = c_skip; s->c1 = c1; s->c2 = c2; s->k0 = c_known + c_skip;
    s->w1_bytes = kop_bytes(c1, s->k0);
    s->w2_bytes = kop_bytes(c2, c1);
    s->stage_a_bytes = kFpTile * 128u;                     // [128 rows][64 bf16]
    s->stage_w_bytes = (uint32_t)max(c1, c2) * 128u;       // [c rows][64 bf16]
    s->a1_bytes = kop_bytes(kFpTile, c1);
    s->smem_bytes = 1024u + 2u * (s->stage_a_bytes + s->stage_w_bytes) + s->a1_bytes + 4u * (c1 + c2) + 64u;
    return s->smem_bytes <= 225u * 1024u;
}

// image = W1 | W2 (K-major swizzled, rows = output channels) | bias1 | bias2
__global__ void fp_pack_weights_kernel(FpTcShape s, const float *__restrict__ w1, const float *__restrict__ b1,
                                       const float *__restrict__ w2, const float *__restrict__ b2,
                                       unsigned char *__restrict__ image)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = s.c1 * s.k0, n2 = s.c2 * s.c1;
    if (i < n1) {
        const int r = i / s.k0, k = i % s.k0;
        *reinterpret_cast<__nv_bfloat16 *>(image + kop_chunk_off(s.c1, s.k0, r, k >> 3) + (k & 7) * 2) =
            __float2bfloat16_rn(w1[i]);
    } else if (i < n1 + n2) {
        const int j = i - n1, r = j / s.c1, k = j % s.c1;
        *reinterpret_cast<__nv_bfloat16 *>(image + s.w1_bytes + kop_chunk_off(s.c2, s.c1, r, k >> 3) + (k & 7) * 2) =
            __float2bfloat16_rn(w2[j]);
    } else if (i < n1 + n2 + s.c1 + s.c2) {
        const int j = i - n1 - n2;
        float *bias = reinterpret_cast<float *>(image + s.w1_bytes + s.w2_bytes);
        bias[j] = j < s.c1 ? (b1 ? b1[j] : 0.f) : (b2 ? b2[j - s.c1] : 0.f);
    }
}

struct FpTcParams {
    FpTcShape s;
    int n, m, tiles_per_scene, ntiles;
    const float *dist2;
    const int *idx;
    const __nv_bfloat16 *known, *skip;
    const unsigned char *image;
    float *out;
    __nv_bfloat16 *out_rows;
};

__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8])
{
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 t = __bfloat1622float2(h[q]);
        f[2 * q] = t.x; f[2 * q + 1] = t.y;
    }
}

__global__ void __launch_bounds__(kFpThreads)
fp_tc_kernel(const FpTcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const FpTcShape &s = p.s;
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char *a1 = base + 2 * (s.stage_a_bytes + s.stage_w_bytes);
    float *bias1 = reinterpret_cast<float *>(a1 + s.a1_bytes);
    float *bias2 = bias1 + s.c1;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(bias2 + s.c2);      // [0],[1]: ring stages, [2]: layer done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 3);

    const int tid = threadIdx.x, warp = tid >> 5;
    {
        const float *bsrc = reinterpret_cast<const float *>(p.image + s.w1_bytes + s.w2_bytes);
        for (int i = tid; i < s.c1 + s.c2; i += kFpThreads) bias1[i] = __ldg(bsrc + i);
        if (tid == 0) {
            tc_mbar_init(smem_u32(mbar), 1);
            tc_mbar_init(smem_u32(mbar + 1), 1);
            tc_mbar_init(smem_u32(mbar + 2), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) tmem_alloc<256>(smem_u32(tmem_slot));
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = *tmem_slot;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc1 = umma_idesc(kFpTile, s.c1), idesc2 = umma_idesc(kFpTile, s.c2);
    const int nslab1 = s.k0 / kSlab, nslab2 = s.c1 / kSlab, nk = s.c_known / kSlab;
    uint32_t uses0 = 0u, uses1 = 0u;  // completed-or-pending uses of each ring stage
    uint32_t done_phase = 0;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int bi = tile / p.tiles_per_scene;
        const int row0 = (tile - bi * p.tiles_per_scene) * kFpTile;
        const int row = row0 + tid;
        const bool valid = row < p.n;
        // interpolation weights of this thread's point (pointnet2_modules.py:399-402)
        float w1 = 0.f, w2 = 0.f, w3 = 0.f;
        const __nv_bfloat16 *f1 = p.known, *f2 = p.known, *f3 = p.known;
        if (valid) {
            const size_t j = ((size_t)bi * p.n + row) * 3;
            const float d1 = __fsqrt_rn(__ldg(p.dist2 + j)), d2 = __fsqrt_rn(__ldg(p.dist2 + j + 1)),
                        d3 = __fsqrt_rn(__ldg(p.dist2 + j + 2));
            const float r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)), r2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f)),
                        r3 = __fdiv_rn(1.0f, __fadd_rn(d3, 1e-8f));
            const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
            w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm); w3 = __fdiv_rn(r3, norm);
            f1 = p.known + ((size_t)bi * p.m + __ldg(p.idx + j)) * s.c_known;
            f2 = p.known + ((size_t)bi * p.m + __ldg(p.idx + j + 1)) * s.c_known;
            f3 = p.known + ((size_t)bi * p.m + __ldg(p.idx + j + 2)) * s.c_known;
        }
        const __nv_bfloat16 *sk = p.skip + ((size_t)bi * p.n + (valid ? row : 0)) * s.c_skip;

        // ---- layer 1: K = c_known + c_skip in slabs of 64, two-stage ring ----
        for (int kb = 0; kb < nslab1 + nslab2; ++kb) {
            const int st = kb & 1;
            const bool layer2 = kb >= nslab1;
            unsigned char *stage_a = base + st * s.stage_a_bytes;
            unsigned char *stage_w = base + 2 * s.stage_a_bytes + st * s.stage_w_bytes;
            const uint32_t uses = st ? uses1 : uses0;
            if (kb == nslab1) {
                // all layer-1 MMAs done -> epilogue 1: + bias, ReLU, bf16 -> A1 (K-major swizzled)
                if (tid == 0) umma_commit(smem_u32(mbar + 2));
                tc_mbar_wait(smem_u32(mbar + 2), done_phase);
                done_phase ^= 1;
                tc_fence_after();
                for (int c0 = 0; c0 < s.c1; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(my_tmem + c0, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int j = q * 8 + h * 2;
                            w[h] = pack_bf16(fmaxf(__uint_as_float(v[j]) + bias1[c0 + j], 0.f),
                                             fmaxf(__uint_as_float(v[j + 1]) + bias1[c0 + j + 1], 0.f));
                        }
                        *reinterpret_cast<uint4 *>(a1 + kop_chunk_off(kFpTile, s.c1, tid, (c0 >> 3) + q)) =
                            make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                tc_fence_before();
            }
            // the MMAs that last read this stage must have finished
            if (uses > 0) tc_mbar_wait(smem_u32(mbar + st), (uses - 1) & 1);
            // stream the weight slab
            {
                const unsigned char *wsrc = layer2 ? p.image + s.w1_bytes + (size_t)(kb - nslab1) * s.c2 * 128u
                                                   : p.image + (size_t)kb * s.c1 * 128u;
                const uint32_t nchunks = (uint32_t)(layer2 ? s.c2 : s.c1) * 8u;
                const uint32_t dst = smem_u32(stage_w);
                for (uint32_t i = tid; i < nchunks; i += kFpThreads) cp_async16(dst + i * 16, wsrc + (size_t)i * 16);
            }
            if (!layer2) {
                const uint32_t adst = smem_u32(stage_a);
                if (kb < nk) {
                    // interpolated channels kb*64 .. +63 of this thread's point
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (valid) {
                            float a[8], b[8], c[8];
                            unpack8(__ldg(reinterpret_cast<const uint4 *>(f1 + kb * kSlab) + ch), a);
                            unpack8(__ldg(reinterpret_cast<const uint4 *>(f2 + kb * kSlab) + ch), b);
                            unpack8(__ldg(reinterpret_cast<const uint4 *>(f3 + kb * kSlab) + ch), c);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                o[q] = pack_bf16(interp3(a[2 * q], w1, b[2 * q], w2, c[2 * q], w3),
                                                 interp3(a[2 * q + 1], w1, b[2 * q + 1], w2, c[2 * q + 1], w3));
                        }
                        *reinterpret_cast<uint4 *>(stage_a + tid * 128 + ((ch ^ (tid & 7)) << 4)) =
                            make_uint4(o[0], o[1], o[2], o[3]);
                    }
                } else {
                    const __nv_bfloat16 *src = sk + (kb - nk) * kSlab;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        if (valid) cp_async16(adst + tid * 128 + ((ch ^ (tid & 7)) << 4), src + ch * 8);
                        else *reinterpret_cast<uint4 *>(stage_a + tid * 128 + ((ch ^ (tid & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
            cp_async_wait_all();
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                const uint32_t elected = elect_one();
                const uint32_t wb = smem_u32(stage_w);
                const int k2 = kb - nslab1;
                // A: this stage's 64-column slab (layer 1) or slab k2 of the layer-1 activations (layer 2)
                const uint64_t da = smem_desc(layer2 ? smem_u32(a1) + (uint32_t)k2 * (kFpTile * 128u) : smem_u32(stage_a),
                                              1024u, kSw128);
                const uint64_t db = smem_desc(wb, 1024u, kSw128);
                const uint32_t idesc = layer2 ? idesc2 : idesc1;
                const uint32_t first = layer2 ? (uint32_t)k2 : (uint32_t)kb;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    if (elected) umma_bf16(tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (first | ks) != 0);
                if (elected) umma_commit(smem_u32(mbar + st));
                __syncwarp();
            }
            if (st) ++uses1; else ++uses0;
        }
        // ---- epilogue 2: + bias, ReLU -> (B, C, n) fp32 and bf16 rows ----
        if (tid == 0) umma_commit(smem_u32(mbar + 2));
        tc_mbar_wait(smem_u32(mbar + 2), done_phase);
        done_phase ^= 1;
        tc_fence_after();
        for (int c0 = 0; c0 < s.c2; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(my_tmem + c0, v);
            if (valid) {
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (c0 + j >= s.c2) break;
                    const float o = fmaxf(__uint_as_float(v[j]) + bias2[c0 + j], 0.f);
                    p.out[((size_t)bi * s.c2 + c0 + j) * p.n + row] = o;       // coalesced over the warp's points
                    v[j] = __float_as_uint(o);
                }
                if (p.out_rows) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) packed[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    uint4 *dst = reinterpret_cast<uint4 *>(p.out_rows + ((size_t)bi * p.n + row) * s.c2 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c0 + q * 8 < s.c2) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_fp_tc_supported(int c_known, int c_skip, int c1, int c2)
{
    FpTcShape s;
    return make_fp_shape(c_known, c_skip, c1, c2, &s) ? 1 : 0;
}

extern "C" size_t pn2_fp_tc_weight_image_bytes(int c_known, int c_skip, int c1, int c2)
{
    FpTcShape s;
    if (!make_fp_shape(c_known, c_skip, c1, c2, &s)) return 0;
    return (size_t)s.w1_bytes + s.w2_bytes + 4u * (c1 + c2);
}

extern "C" int pn2_fp_tc_pack_weights(int c_known, int c_skip, int c1, int c2, const float *w1, const float *b1,
                                      const float *w2, const float *b2, void *image, pn2_stream_t stream)
{
    FpTcShape s;
    if (!make_fp_shape(c_known, c_skip, c1, c2, &s) || !w1 || !w2 || !image) return PN2_ERR_INVALID_ARGUMENT;
    const int total = s.c1 * s.k0 + s.c2 * s.c1 + s.c1 + s.c2;
    fp_pack_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(s, w1, b1, w2, b2,
                                                                                static_cast<unsigned char *>(image));
    PN2_LAUNCH_CHECK("fp_tc_pack_weights");
    return PN2_OK;
}

extern "C" int pn2_fp_tc_forward(int b, int n, int m, int c_known, int c_skip, int c1, int c2, const float *dist2,
                                 const int *idx, const void *known_rows, const void *skip_rows,
                                 const void *weight_image, float *out, void *out_rows, pn2_stream_t stream)
{
    FpTcParams p;
    if (b < 0 || n < 0 || m < 1 || !make_fp_shape(c_known, c_skip, c1, c2, &p.s)) return PN2_ERR_INVALID_ARGUMENT;
    if (b == 0 || n == 0) return PN2_OK;
    if (!dist2 || !idx || !known_rows || (c_skip > 0 && !skip_rows) || !weight_image || !out) return PN2_ERR_INVALID_ARGUMENT;
    p.n = n; p.m = m;
    p.tiles_per_scene = ceil_div(n, kFpTile);
    p.ntiles = b * p.tiles_per_scene;
    p.dist2 = dist2; p.idx = idx;
    p.known = static_cast<const __nv_bfloat16 *>(known_rows);
    p.skip = static_cast<const __nv_bfloat16 *>(c_skip > 0 ? skip_rows : known_rows);
    p.image = static_cast<const unsigned char *>(weight_image);
    p.out = out;
    p.out_rows = static_cast<__nv_bfloat16 *>(out_rows);
    const int sms = stream_sm_count(as_stream(stream));
    PN2_CUDA_TRY(cudaFuncSetAttribute(fp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.s.smem_bytes));
    fp_tc_kernel<<<min(p.ntiles, sms), kFpThreads, p.s.smem_bytes, as_stream(stream)>>>(p);
    PN2_LAUNCH_CHECK("fp_tc_forward");
    return PN2_OK;
}
