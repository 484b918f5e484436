// tc_common.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers and the K-major operand layout shared by the
// tensor-core kernels (sa_tc.cu, fp_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace pn2 {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// -DPN2_TC_MBAR_HINT=0x989680 builds try_wait with a suspend-time hint (a waiting warp sleeps in hardware until the phase
// completes instead of re-issuing the probe + branch).  Measured neutral on every kernel of the step (SA1 59.4 us both
// ways, scripts/gpu_r2_hint.sh), so the plain form stays the default.
#ifndef PN2_TC_MBAR_HINT
#define PN2_TC_MBAR_HINT 0
#endif
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
#if PN2_TC_MBAR_HINT
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)PN2_TC_MBAR_HINT) : "memory");
#else
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot)
{
    const uint32_t cols = COLS;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    const uint32_t cols = COLS;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- K-major operand tiles ------------------------------------------------------------------------
// A `rows` x K bf16 operand (K multiple of 16) is stored as K/64 tiles of [rows][64] in the canonical
// 128-byte-swizzle K-major layout (row pitch 128 B, 8-row groups of 1024 B, 16-byte chunk index XORed
// with row%8) followed by (K%64)/16 tiles of [rows][16] in the 32-byte-swizzle layout (row pitch 32 B,
// 8-row groups of 256 B, chunk index XORed with (row/4)%2).  One tcgen05.mma consumes K = 16.
__host__ __device__ inline uint32_t kop_bytes(int rows, int K) { return (uint32_t)rows * (128u * (K / 64) + 32u * ((K % 64) / 16)); }

// byte offset of the 16-byte chunk that holds elements [8*ch, 8*ch + 8) of row r
__host__ __device__ inline uint32_t kop_chunk_off(int rows, int K, int r, int ch)
{
    const int nfull = K / 64;
    if (ch < nfull * 8) return (uint32_t)(ch >> 3) * rows * 128u + r * 128u + (((ch & 7) ^ (r & 7)) << 4);
    const int t = (ch - nfull * 8) >> 1;
    return (uint32_t)nfull * rows * 128u + (uint32_t)t * rows * 32u + r * 32u + ((((ch & 1) ^ ((r >> 2) & 1))) << 4);
}

// shared-memory matrix descriptor (SM100 format: version 1; LBO unused for swizzled K-major)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t kSw128 = 2, kSw32 = 6;

// descriptor of K-step ks (16 elements) of an operand whose first row is r0 (multiple of 8)
__device__ __forceinline__ uint64_t kop_desc(uint32_t base, int rows, int K, int ks, int r0)
{
    const int nfull4 = (K / 64) * 4;
    if (ks < nfull4)
        return smem_desc(base + (uint32_t)(ks >> 2) * rows * 128u + r0 * 128u + (ks & 3) * 32u, 1024u, kSw128);
    return smem_desc(base + (uint32_t)(K / 64) * rows * 128u + (uint32_t)(ks - nfull4) * rows * 32u + r0 * 32u, 256u,
                     kSw32);
}

// one elected lane of a converged warp (the other lanes get 0)
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t e;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(e));
    return e;
}

// D[tmem_d] = A (a_rows x K at a_base, first row a_r0) * B (b_rows x K at b_base, first row b_r0)^T, issued by
// the elected lane of a converged warp.  Called by ALL lanes: the descriptors are computed warp-uniformly
// (uniform registers) and only the tcgen05.mma itself is predicated -- measured on B200 this issues an MMA
// every ~45 cycles instead of ~70 when a single thread builds the descriptors in vector registers.
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_base, int a_rows, int a_r0, uint32_t b_base,
                                           int b_rows, int b_r0, int K, uint32_t idesc, uint32_t elected)
{
    const int nfull = K >> 6, ntail = (K & 63) >> 4;
    const uint64_t da = smem_desc(a_base + (uint32_t)a_r0 * 128u, 1024u, kSw128);
    const uint64_t db = smem_desc(b_base + (uint32_t)b_r0 * 128u, 1024u, kSw128);
    const uint32_t a_step = (uint32_t)a_rows * 8u, b_step = (uint32_t)b_rows * 8u;   // 128-byte rows, 16-byte units
    for (int kb = 0; kb < nfull; ++kb) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (elected)
                umma_bf16(tmem_d, da + (uint64_t)(kb * a_step + j * 2), db + (uint64_t)(kb * b_step + j * 2), idesc,
                          (uint32_t)((kb | j) != 0));
    }
    const uint32_t a_tail = a_base + (uint32_t)nfull * a_rows * 128u + (uint32_t)a_r0 * 32u;
    const uint32_t b_tail = b_base + (uint32_t)nfull * b_rows * 128u + (uint32_t)b_r0 * 32u;
    for (int t = 0; t < ntail; ++t)
        if (elected)
            umma_bf16(tmem_d, smem_desc(a_tail + (uint32_t)t * a_rows * 32u, 256u, kSw32),
                      smem_desc(b_tail + (uint32_t)t * b_rows * 32u, 256u, kSw32), idesc, (uint32_t)((nfull | t) != 0));
}

// instruction descriptor: D f32, A/B bf16, both K-major, N at [17,23) in units of 8, M at [24,29) in units of 16
__host__ __device__ inline uint32_t umma_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// issue only; pair with tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two floats -> packed bf16x2 with ReLU folded into the conversion (lo in the low half)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}


static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace pn2
