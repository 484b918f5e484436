// lib.cu -- library-level entry points of libpn2_b200: version, error text, device check.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace pn2 {

static thread_local char g_cuda_error[512] = "";

void set_cuda_error(cudaError_t e, const char *where)
{
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", where, cudaGetErrorString(e),
             cudaGetErrorName(e));
}

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

}  // namespace pn2

extern "C" unsigned long long pn2_launch_count(void) { return __atomic_load_n(&pn2::g_launches, __ATOMIC_RELAXED); }

extern "C" int pn2_version(void) { return 0 * 10000 + 1 * 100 + 0; }

extern "C" const char *pn2_error_string(int code)
{
    switch (code) {
    case PN2_OK: return "ok";
    case PN2_ERR_INVALID_ARGUMENT: return "invalid argument (dims, null pointer or unsupported shape)";
    case PN2_ERR_CUDA: return "CUDA call or kernel launch failed (see pn2_last_cuda_error)";
    case PN2_ERR_WORKSPACE: return "workspace missing or too small";
    case PN2_ERR_UNSUPPORTED_DEVICE: return "current device is not compute capability 10.x (sm_100a build)";
    default: return "unknown error code";
    }
}

extern "C" const char *pn2_last_cuda_error(void) { return pn2::g_cuda_error; }

extern "C" int pn2_device_check(void)
{
    int dev = 0;
    PN2_CUDA_TRY(cudaGetDevice(&dev));
    int major = 0;
    PN2_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    return major == 10 ? PN2_OK : PN2_ERR_UNSUPPORTED_DEVICE;
}
