// lin_tc.cu -- per-point half of a set-abstraction layer's first 1x1 convolution, on tcgen05.
//
// The reference groups first and convolves afterwards (pointnet2_utils.py:348-359 -> pytorch_utils.py:11-36):
// layer 1 of an SA module's SharedMLP multiplies the (3 + C)-channel vector of every (centre, neighbour)
// pair by W1.  Before the ReLU that layer is linear, and only the three xyz channels depend on the centre:
//     W1 [ (p_nb - p_c)/r ; f_nb ]  =  W1[:, :3] (p_nb - p_c)/r  +  W1[:, 3:] f_nb
// so the feature half  P[n] = W1[:, 3:] f_n  is a property of the POINT and is computed here once per
// point instead of once per pair (SA1: 40 000 points vs 131 072 pairs per scene, with C = 129 -> 64
// channels; SA3/SA4: 8x fewer rows).  The fused SA kernel (sa_tc.cu) then gathers the 64/128-wide P rows
// (128/256 bytes: whole cache lines) instead of the (3+C)-wide feature rows, and its layer 1 shrinks to the
// xyz/bias columns plus an identity block.  P is stored as bf16 (same rounding unit as every other
// activation of the bf16 arm).
//
// Kernel: a plain row GEMM  P (rows x C1) = X (rows x K) W^T, 128-row tiles, persistent CTAs:
//   warps 0-7   load the tile (fp32 rows straight out of point_clouds, converted to bf16 in registers, or
//               bf16 rows of a previous layer) into the K-major 128B-swizzled A operand, double buffered
//   warp 8      issues the tcgen05.mma chain, accumulator double buffered in TMEM
//   warps 12-15 epilogue: TMEM -> bf16 -> one contiguous row per thread
// For fp32 input the A operand is the WHOLE raw row (xyz included, so that every 16-byte load is aligned);
// the weight image has zero columns under xyz.  HBM-bound: it reads the input once.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pn2 {

constexpr int kLinTile = 128;
constexpr int kLinLoaders = 256;         // warps 0-7
constexpr int kLinWarpMma = 8;
constexpr int kLinEpiWarp0 = 12;         // warps 12-15 (TMEM lane quarter = warp % 4)
constexpr int kLinThreads = 512;

struct LinParams {
    long long rows;
    int kin;                // A columns taken from the input row (f32: skip + c, multiple of 4; bf16: row_elems, multiple of 8)
    int K;                  // round_up(kin, 16)
    int ld;                 // input row pitch in elements
    int ntiles;
    const void *x;
    const unsigned char *image;
    __nv_bfloat16 *out;
};

__global__ void lin_pack_weights_kernel(int kin, int K, int skip, int c, int c1, const float *__restrict__ w, int w_ld,
                                        int w_col0, unsigned char *__restrict__ image)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c1 * K) return;
    const int j = i / K, e = i - j * K;
    float v = 0.f;
    if (e >= skip && e < skip + c && e < kin) v = w[(size_t)j * w_ld + w_col0 + (e - skip)];
    *reinterpret_cast<__nv_bfloat16 *>(image + kop_chunk_off(c1, K, j, e >> 3) + (e & 7) * 2) = __float2bfloat16_rn(v);
}

template <int C1, bool BF16_IN>
__global__ void __launch_bounds__(kLinThreads, 1)
lin_tc_kernel(const LinParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int K = p.K;
    const uint32_t w_bytes = (kop_bytes(C1, K) + 1023u) & ~1023u, a_bytes = (kop_bytes(kLinTile, K) + 1023u) & ~1023u;
    unsigned char *ws = base, *as = base + w_bytes;                          // A buffers: as, as + a_bytes
    uint64_t *mbar = reinterpret_cast<uint64_t *>(as + 2 * a_bytes);         // a_full[2], a_empty[2], d_full[2], d_empty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 8);
    const uint32_t bar = smem_u32(mbar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.image);
        for (uint32_t i = tid; i < kop_bytes(C1, K) / 16; i += kLinThreads) cp_async16(smem_u32(ws) + i * 16, src + i);
        // columns [kin, K) of both A buffers stay zero for the whole kernel
        const int padq = (K - p.kin) / 4;                                    // 8-byte pieces of 4 bf16
        for (int i = tid; i < 2 * kLinTile * padq; i += kLinThreads) {
            const int b = i / (kLinTile * padq), r = (i / padq) % kLinTile, e = p.kin + (i % padq) * 4;
            *reinterpret_cast<uint2 *>(as + b * a_bytes + kop_chunk_off(kLinTile, K, r, e >> 3) + (e & 7) * 2) = make_uint2(0u, 0u);
        }
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) {
                tc_mbar_init(bar + 8u * i, kLinLoaders / 32);      // a_full: one arrival per loader warp
                tc_mbar_init(bar + 8u * (2 + i), 1);               // a_empty: tcgen05.commit
                tc_mbar_init(bar + 8u * (4 + i), 1);               // d_full: tcgen05.commit
                tc_mbar_init(bar + 8u * (6 + i), 4);               // d_empty: one arrival per epilogue warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) tmem_alloc<2 * C1 < 32 ? 32 : 2 * C1>(smem_u32(tmem_slot));
        cp_async_wait_all();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = *tmem_slot;
    const int nt = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp < kLinLoaders / 32) {
        // ===== loaders =====
        // All of a thread's 16-byte loads of a tile are issued before the first one is used (one memory round trip
        // per tile and ~70 KB in flight per SM); (row, column) advance incrementally, no division in the loop.
        constexpr int U = 17;
        const int per_row = BF16_IN ? p.kin / 8 : p.kin / 4;                 // 16-byte pieces per row
        const int total = kLinTile * per_row;
        const int step_r = kLinLoaders / per_row, step_c = kLinLoaders - step_r * per_row;
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            const long long row0 = (long long)(blockIdx.x + it * gridDim.x) * kLinTile;
            unsigned char *a = as + st * a_bytes;
            for (int f0 = tid; f0 < total; f0 += kLinLoaders * U) {
                uint4 v[U];
                int r = f0 / per_row, c = f0 - r * per_row;
                {
                    int rr = r, cc = c;
#pragma unroll
                    for (int q = 0; q < U; ++q) {
                        v[q] = make_uint4(0u, 0u, 0u, 0u);
                        if (rr < kLinTile && row0 + rr < p.rows) {
                            if (BF16_IN)
                                v[q] = __ldg(reinterpret_cast<const uint4 *>(static_cast<const __nv_bfloat16 *>(p.x) +
                                                                             (size_t)(row0 + rr) * p.ld + cc * 8));
                            else
                                v[q] = __ldg(reinterpret_cast<const uint4 *>(static_cast<const float *>(p.x) +
                                                                             (size_t)(row0 + rr) * p.ld + cc * 4));
                        }
                        rr += step_r; cc += step_c;
                        if (cc >= per_row) { cc -= per_row; ++rr; }
                    }
                }
                if (f0 == tid && u > 0) tc_mbar_wait(bar + 8u * (2 + st), (u - 1) & 1);   // buffer free (loads already in flight)
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    if (r < kLinTile) {
                        if (BF16_IN) {
                            *reinterpret_cast<uint4 *>(a + kop_chunk_off(kLinTile, K, r, c)) = v[q];
                        } else {
                            *reinterpret_cast<uint2 *>(a + kop_chunk_off(kLinTile, K, r, c >> 1) + (c & 1) * 8) =
                                make_uint2(pack_bf16(__uint_as_float(v[q].x), __uint_as_float(v[q].y)),
                                           pack_bf16(__uint_as_float(v[q].z), __uint_as_float(v[q].w)));
                        }
                    }
                    r += step_r; c += step_c;
                    if (c >= per_row) { c -= per_row; ++r; }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + 8u * st) : "memory");
            }
        }
    } else if (warp == kLinWarpMma) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc(kLinTile, C1);
        const uint32_t elected = elect_one();
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            tc_mbar_wait(bar + 8u * st, u & 1);                              // tile loaded
            if (u > 0) tc_mbar_wait(bar + 8u * (6 + st), (u - 1) & 1);       // accumulator drained
            fence_proxy_async();
            tc_fence_after();
            issue_gemm(tmem + st * C1, smem_u32(as + st * a_bytes), kLinTile, 0, smem_u32(ws), C1, 0, K, idesc, elected);
            if (elected) {
                umma_commit(bar + 8u * (4 + st));
                umma_commit(bar + 8u * (2 + st));
            }
            __syncwarp();
        }
    } else if (warp >= kLinEpiWarp0) {
        // ===== epilogue: one output row per thread =====
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t my_tmem = tmem + ((uint32_t)(quarter * 32) << 16);
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            const long long grow = (long long)(blockIdx.x + it * gridDim.x) * kLinTile + row;
            tc_mbar_wait(bar + 8u * (4 + st), u & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < C1; c0 += 64) {
                uint32_t va[32], vb[32];
                tmem_ld32_issue(my_tmem + st * C1 + c0, va);
                tmem_ld32_issue(my_tmem + st * C1 + c0 + 32, vb);
                tmem_ld_wait();
                if (grow < p.rows) {
                    uint4 *dst = reinterpret_cast<uint4 *>(p.out + (size_t)grow * C1 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_uint4(pack_bf16(__uint_as_float(va[8 * q]), __uint_as_float(va[8 * q + 1])),
                                            pack_bf16(__uint_as_float(va[8 * q + 2]), __uint_as_float(va[8 * q + 3])),
                                            pack_bf16(__uint_as_float(va[8 * q + 4]), __uint_as_float(va[8 * q + 5])),
                                            pack_bf16(__uint_as_float(va[8 * q + 6]), __uint_as_float(va[8 * q + 7])));
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[4 + q] = make_uint4(pack_bf16(__uint_as_float(vb[8 * q]), __uint_as_float(vb[8 * q + 1])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 2]), __uint_as_float(vb[8 * q + 3])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 4]), __uint_as_float(vb[8 * q + 5])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 6]), __uint_as_float(vb[8 * q + 7])));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + 8u * (6 + st)) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<2 * C1 < 32 ? 32 : 2 * C1>(tmem);
}

// ---- fp32 rows that are contiguous in memory (ld == kin): the raw tile is one 1-D bulk copy (TMA) ----------
// The register-staged loader above leaves HBM idle while it converts and the converters idle while they wait
// (ncu: 35 % of DRAM peak).  Here one thread streams whole tiles (128 rows x kin floats, up to 67.6 KB) into a
// double-buffered raw area with cp.async.bulk, two tiles ahead of the converters, so the DRAM queue never drains:
//   warp 9      cp.async.bulk global -> shared, completion on a transaction mbarrier
//   warps 0-7   raw fp32 (ld.shared.v4) -> bf16 -> K-major swizzled A operand
//   warp 8      tcgen05.mma chain, accumulator double buffered in TMEM
//   warps 12-15 epilogue
constexpr int kLinWarpTma = 9;

template <int C1>
__global__ void __launch_bounds__(kLinThreads, 1)
lin_tc_tma_kernel(const LinParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    unsigned char *base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int K = p.K;
    const uint32_t w_bytes = (kop_bytes(C1, K) + 1023u) & ~1023u, a_bytes = (kop_bytes(kLinTile, K) + 1023u) & ~1023u;
    const uint32_t raw_bytes = ((uint32_t)kLinTile * p.kin * 4u + 127u) & ~127u;
    unsigned char *ws = base, *as = base + w_bytes, *rawb = as + a_bytes;    // raw buffers: rawb, rawb + raw_bytes
    uint64_t *mbar = reinterpret_cast<uint64_t *>(rawb + 2 * raw_bytes);
    // raw_full[2], raw_empty[2], a_full, a_empty, d_full[2], d_empty[2]
    const uint32_t bar = smem_u32(mbar);
    const uint32_t b_raw_full = bar, b_raw_empty = bar + 16, b_a_full = bar + 32, b_a_empty = bar + 40, b_d_full = bar + 48,
                   b_d_empty = bar + 64;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 10);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.image);
        for (uint32_t i = tid; i < kop_bytes(C1, K) / 16; i += kLinThreads) cp_async16(smem_u32(ws) + i * 16, src + i);
        const int padq = (K - p.kin) / 4;
        for (int i = tid; i < kLinTile * padq; i += kLinThreads) {
            const int r = i / padq, e = p.kin + (i % padq) * 4;
            *reinterpret_cast<uint2 *>(as + kop_chunk_off(kLinTile, K, r, e >> 3) + (e & 7) * 2) = make_uint2(0u, 0u);
        }
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) {
                tc_mbar_init(b_raw_full + 8u * i, 1);                  // expect_tx arrival + the copy's bytes
                tc_mbar_init(b_raw_empty + 8u * i, kLinLoaders / 32);  // one arrival per converter warp
                tc_mbar_init(b_d_full + 8u * i, 1);
                tc_mbar_init(b_d_empty + 8u * i, 4);
            }
            tc_mbar_init(b_a_full, kLinLoaders / 32);
            tc_mbar_init(b_a_empty, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) tmem_alloc<2 * C1 < 32 ? 32 : 2 * C1>(smem_u32(tmem_slot));
        cp_async_wait_all();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = *tmem_slot;
    const int nt = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto tile_rows = [&](int it) -> int {
        const long long row0 = (long long)(blockIdx.x + it * gridDim.x) * kLinTile;
        const long long left = p.rows - row0;
        return left < kLinTile ? (int)left : kLinTile;
    };

    if (warp == kLinWarpTma) {
        if (elect_one()) {
            const float *x = static_cast<const float *>(p.x);
            for (int it = 0; it < nt; ++it) {
                const int st = it & 1, u = it >> 1;
                if (u > 0) tc_mbar_wait(b_raw_empty + 8u * st, (u - 1) & 1);
                const long long row0 = (long long)(blockIdx.x + it * gridDim.x) * kLinTile;
                const uint32_t bytes = (uint32_t)tile_rows(it) * (uint32_t)p.kin * 4u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_raw_full + 8u * st), "r"(bytes)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                    ::"r"(smem_u32(rawb + st * raw_bytes)), "l"(x + (size_t)row0 * p.kin), "r"(bytes),
                      "r"(b_raw_full + 8u * st)
                    : "memory");
            }
        }
        __syncwarp();
    } else if (warp < kLinLoaders / 32) {
        // ===== converters =====
        const int per_row = p.kin / 4, total = kLinTile * per_row;
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            const int nrows = tile_rows(it);
            const unsigned char *rsrc = rawb + st * raw_bytes;
            tc_mbar_wait(b_raw_full + 8u * st, u & 1);
            if (it > 0) tc_mbar_wait(b_a_empty, (it - 1) & 1);
            int r = tid / per_row, c = tid - r * per_row;
            const int step_r = kLinLoaders / per_row, step_c = kLinLoaders - step_r * per_row;
            for (int f = tid; f < total; f += kLinLoaders) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (r < nrows) v = *reinterpret_cast<const uint4 *>(rsrc + ((size_t)r * per_row + c) * 16);
                *reinterpret_cast<uint2 *>(as + kop_chunk_off(kLinTile, K, r, c >> 1) + (c & 1) * 8) =
                    make_uint2(pack_bf16(__uint_as_float(v.x), __uint_as_float(v.y)),
                               pack_bf16(__uint_as_float(v.z), __uint_as_float(v.w)));
                r += step_r; c += step_c;
                if (c >= per_row) { c -= per_row; ++r; }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_a_full) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_raw_empty + 8u * st) : "memory");
            }
        }
    } else if (warp == kLinWarpMma) {
        const uint32_t idesc = umma_idesc(kLinTile, C1);
        const uint32_t elected = elect_one();
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            tc_mbar_wait(b_a_full, it & 1);
            if (u > 0) tc_mbar_wait(b_d_empty + 8u * st, (u - 1) & 1);
            fence_proxy_async();
            tc_fence_after();
            issue_gemm(tmem + st * C1, smem_u32(as), kLinTile, 0, smem_u32(ws), C1, 0, K, idesc, elected);
            if (elected) {
                umma_commit(b_d_full + 8u * st);
                umma_commit(b_a_empty);
            }
            __syncwarp();
        }
    } else if (warp >= kLinEpiWarp0) {
        const int quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t my_tmem = tmem + ((uint32_t)(quarter * 32) << 16);
        for (int it = 0; it < nt; ++it) {
            const int st = it & 1, u = it >> 1;
            const long long grow = (long long)(blockIdx.x + it * gridDim.x) * kLinTile + row;
            tc_mbar_wait(b_d_full + 8u * st, u & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < C1; c0 += 64) {
                uint32_t va[32], vb[32];
                tmem_ld32_issue(my_tmem + st * C1 + c0, va);
                tmem_ld32_issue(my_tmem + st * C1 + c0 + 32, vb);
                tmem_ld_wait();
                if (grow < p.rows) {
                    uint4 *dst = reinterpret_cast<uint4 *>(p.out + (size_t)grow * C1 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_uint4(pack_bf16(__uint_as_float(va[8 * q]), __uint_as_float(va[8 * q + 1])),
                                            pack_bf16(__uint_as_float(va[8 * q + 2]), __uint_as_float(va[8 * q + 3])),
                                            pack_bf16(__uint_as_float(va[8 * q + 4]), __uint_as_float(va[8 * q + 5])),
                                            pack_bf16(__uint_as_float(va[8 * q + 6]), __uint_as_float(va[8 * q + 7])));
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[4 + q] = make_uint4(pack_bf16(__uint_as_float(vb[8 * q]), __uint_as_float(vb[8 * q + 1])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 2]), __uint_as_float(vb[8 * q + 3])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 4]), __uint_as_float(vb[8 * q + 5])),
                                                pack_bf16(__uint_as_float(vb[8 * q + 6]), __uint_as_float(vb[8 * q + 7])));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b_d_empty + 8u * st) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<2 * C1 < 32 ? 32 : 2 * C1>(tmem);
}

static uint32_t lin_tma_smem_bytes(int kin, int K, int c1)
{
    return 1024u + ((kop_bytes(c1, K) + 1023u) & ~1023u) + ((kop_bytes(kLinTile, K) + 1023u) & ~1023u) +
           2u * (((uint32_t)kLinTile * kin * 4u + 127u) & ~127u) + 10u * 8u + 16u;
}

template <int C1>
static int launch_lin_tma(const LinParams &p, cudaStream_t stream)
{
    const int sms = stream_sm_count(stream);
    const uint32_t smem = lin_tma_smem_bytes(p.kin, p.K, C1);
    auto kern = lin_tc_tma_kernel<C1>;
    PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<min(p.ntiles, sms), kLinThreads, smem, stream>>>(p);
    PN2_LAUNCH_CHECK("lin_tc_forward(bulk)");
    return PN2_OK;
}

static uint32_t lin_smem_bytes(int K, int c1)
{
    return 1024u + ((kop_bytes(c1, K) + 1023u) & ~1023u) + 2u * ((kop_bytes(kLinTile, K) + 1023u) & ~1023u) + 8u * 8u + 16u;
}

template <int C1, bool BF16_IN>
static int launch_lin(const LinParams &p, cudaStream_t stream)
{
    const int sms = stream_sm_count(stream);
    const uint32_t smem = lin_smem_bytes(p.K, C1);
    auto kern = lin_tc_kernel<C1, BF16_IN>;
    PN2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<min(p.ntiles, sms), kLinThreads, smem, stream>>>(p);
    PN2_LAUNCH_CHECK("lin_tc_forward");
    return PN2_OK;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_lin_tc_supported(int kin, int c1, int x_is_bf16)
{
    if (c1 != 64 && c1 != 128) return 0;
    if (kin < 8 || kin > 512) return 0;
    if (x_is_bf16 ? (kin % 8 != 0) : (kin % 4 != 0)) return 0;
    return lin_smem_bytes(rup(kin, 16), c1) <= 227u * 1024u ? 1 : 0;
}

extern "C" size_t pn2_lin_tc_weight_image_bytes(int kin, int c1)
{
    if (kin < 1 || c1 < 16) return 0;
    return kop_bytes(c1, rup(kin, 16));
}

extern "C" int pn2_lin_tc_pack_weights(int kin, int skip, int c, int c1, const float *w, int w_ld, int w_col0,
                                       void *image, pn2_stream_t stream)
{
    if (kin < 1 || skip < 0 || c < 1 || skip + c > kin || c1 < 16 || !w || !image || w_ld < w_col0 + c)
        return PN2_ERR_INVALID_ARGUMENT;
    const int K = rup(kin, 16);
    lin_pack_weights_kernel<<<ceil_div((long long)c1 * K, 256), 256, 0, as_stream(stream)>>>(
        kin, K, skip, c, c1, w, w_ld, w_col0, static_cast<unsigned char *>(image));
    PN2_LAUNCH_CHECK("lin_tc_pack_weights");
    return PN2_OK;
}

extern "C" int pn2_lin_tc_forward(long long rows, int kin, int c1, const void *x, int x_is_bf16, int ld,
                                  const void *weight_image, void *out, pn2_stream_t stream)
{
    if (rows < 0 || !pn2_lin_tc_supported(kin, c1, x_is_bf16) || ld < kin) return PN2_ERR_INVALID_ARGUMENT;
    if (rows == 0) return PN2_OK;
    if (!x || !weight_image || !out) return PN2_ERR_INVALID_ARGUMENT;
    if (x_is_bf16 ? (ld % 8 != 0) : (ld % 4 != 0)) return PN2_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0) return PN2_ERR_INVALID_ARGUMENT;
    LinParams p;
    p.rows = rows; p.kin = kin; p.K = rup(kin, 16); p.ld = ld;
    p.ntiles = (int)((rows + kLinTile - 1) / kLinTile);
    p.x = x;
    p.image = static_cast<const unsigned char *>(weight_image);
    p.out = static_cast<__nv_bfloat16 *>(out);
    // contiguous fp32 rows: stream the raw tiles with bulk copies when the staging buffers fit
    static const bool use_bulk = [] { const char *e = getenv("PN2_LIN_BULK"); return !e || atoi(e) != 0; }();
    if (use_bulk && !x_is_bf16 && ld == kin && lin_tma_smem_bytes(kin, p.K, c1) <= 227u * 1024u)
        return c1 == 64 ? launch_lin_tma<64>(p, as_stream(stream)) : launch_lin_tma<128>(p, as_stream(stream));
    if (c1 == 64) return x_is_bf16 ? launch_lin<64, true>(p, as_stream(stream)) : launch_lin<64, false>(p, as_stream(stream));
    return x_is_bf16 ? launch_lin<128, true>(p, as_stream(stream)) : launch_lin<128, false>(p, as_stream(stream));
}
