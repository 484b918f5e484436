// common.cuh -- shared helpers for the sm_100a kernels of libpn2_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/pn2_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpn2_b200 is written for sm_100a only"
#endif

namespace pn2 {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; persistent grids are sized from this

// ---- error plumbing -------------------------------------------------------
void set_cuda_error(cudaError_t e, const char *where);
void count_launches(int n);
int stream_sm_count(cudaStream_t stream);   // sm_partition.cu: SMs of the stream's partition, else of the device   // pn2_launch_count(): kernels this library has launched (bench.py's gpu_launches)

#define PN2_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            ::pn2::set_cuda_error(_e, #expr);                \
            return PN2_ERR_CUDA;                             \
        }                                                    \
    } while (0)

#define PN2_LAUNCH_CHECK(name)                               \
    do {                                                     \
        ::pn2::count_launches(1);                            \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) {                             \
            ::pn2::set_cuda_error(_e, name);                 \
            return PN2_ERR_CUDA;                             \
        }                                                    \
    } while (0)

static inline cudaStream_t as_stream(pn2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- the one distance recipe every index-producing kernel shares ----------
// nvcc contracts the reference's (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f) into
// FMUL, FFMA, FFMA with the x term first; written explicitly so
// that no compiler version can reassociate it.  `a` is the minuend.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// three_nn's variant: in the reference build nvcc multiplies the y term first and fuses x, then z
// (SASS of the reference's three_nn_kernel; pinned by tests/golden/ref_cuda_ops.npz).
__device__ __forceinline__ float sqdist3_yxz(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// three_interpolate's p1*w1 + p2*w2 + p3*w3: FMUL on the second term, then the first, then the
// third (same build, same pinning).
__device__ __forceinline__ float interp3(float p1, float w1, float p2, float w2, float p3, float w3)
{
    return __fmaf_rn(p3, w3, __fmaf_rn(p1, w1, __fmul_rn(p2, w2)));
}

__device__ __forceinline__ float sqnorm3(float x, float y, float z)
{
    return __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace pn2
