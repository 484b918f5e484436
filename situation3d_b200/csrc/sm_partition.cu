// sm_partition.cu -- SM partitions (CUDA green contexts) for callers that keep several batches in flight.
//
// The sampling chain of a batch (fps.cu) is latency-bound and occupies half an SM per CTA for ~2 ms, while
// the fused MLP kernels are persistent, take a whole SM per CTA and divide their tiles statically.  When
// both run on the same SMs the hardware scheduler spreads the sampling CTAs over (nearly) all SMs, and a
// fused kernel's CTAs wait for SMs that never become entirely free.  A partition gives the sampling streams
// their own group of SMs (where their CTAs pack two per SM) and everything else the remaining SMs.
// Host code only: driver API (cuGreenCtx*), no kernels.
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>

namespace pn2 {

struct SmPartition {
    CUgreenCtx ctx[2];
    int sms[2];
};

static std::mutex g_stream_mu;
static std::unordered_map<cudaStream_t, int> g_stream_sms;   // streams created here -> SMs of their group

// The driver entry points are resolved at run time (cudaGetDriverEntryPoint): the library must load on machines
// without libcuda (the CPU-only build container).
template <typename F>
static bool driver_fn(const char *name, F *fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p)
        return false;
    *fn = reinterpret_cast<F>(p);
    return true;
}
struct DriverApi {
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    CUresult (*CtxGetDevice)(CUdevice *) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *,
                                          unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
    DriverApi()
    {
        ok = driver_fn("cuGetErrorString", &GetErrorString) && driver_fn("cuCtxGetDevice", &CtxGetDevice) &&
             driver_fn("cuDeviceGetDevResource", &DeviceGetDevResource) &&
             driver_fn("cuDevSmResourceSplitByCount", &DevSmResourceSplitByCount) &&
             driver_fn("cuDevResourceGenerateDesc", &DevResourceGenerateDesc) &&
             driver_fn("cuGreenCtxCreate", &GreenCtxCreate) && driver_fn("cuGreenCtxStreamCreate", &GreenCtxStreamCreate);
    }
};
static const DriverApi &driver()
{
    static DriverApi api;
    return api;
}

#define PN2_CU_TRY(expr)                                                          \
    do {                                                                          \
        CUresult _r = (expr);                                                     \
        if (_r != CUDA_SUCCESS) {                                                 \
            const char *_s = nullptr;                                             \
            driver().GetErrorString(_r, &_s);                                            \
            ::pn2::set_cuda_error(cudaErrorUnknown, _s ? _s : #expr);             \
            return PN2_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

// SMs a persistent kernel launched on `stream` can occupy (its partition's, else the device's)
int stream_sm_count(cudaStream_t stream)
{
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        auto it = g_stream_sms.find(stream);
        if (it != g_stream_sms.end()) return it->second;
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_sm_partition_create(int sms_first, void **handle)
{
    if (!handle || sms_first < 1) return PN2_ERR_INVALID_ARGUMENT;
    PN2_CUDA_TRY(cudaFree(0));                                   // make sure the primary context exists
    const DriverApi &cu = driver();
    if (!cu.ok) { set_cuda_error(cudaErrorNotSupported, "green-context driver entry points"); return PN2_ERR_CUDA; }
    CUdevice dev;
    PN2_CU_TRY(cu.CtxGetDevice(&dev));
    CUdevResource all, first, rest;
    PN2_CU_TRY(cu.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    if ((unsigned)sms_first >= all.sm.smCount) return PN2_ERR_INVALID_ARGUMENT;
    unsigned groups = 1;
    PN2_CU_TRY(cu.DevSmResourceSplitByCount(&first, &groups, &all, &rest, 0, (unsigned)sms_first));
    if (groups != 1 || rest.sm.smCount == 0) return PN2_ERR_INVALID_ARGUMENT;
    SmPartition *p = new SmPartition();
    CUdevResource *res[2] = {&first, &rest};
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        PN2_CU_TRY(cu.DevResourceGenerateDesc(&desc, res[i], 1));
        PN2_CU_TRY(cu.GreenCtxCreate(&p->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
        p->sms[i] = (int)res[i]->sm.smCount;
    }
    *handle = p;
    return PN2_OK;
}

extern "C" int pn2_sm_partition_sms(void *handle, int which)
{
    if (!handle || which < 0 || which > 1) return 0;
    return static_cast<SmPartition *>(handle)->sms[which];
}

extern "C" int pn2_sm_partition_stream_create(void *handle, int which, void **stream)
{
    if (!handle || which < 0 || which > 1 || !stream) return PN2_ERR_INVALID_ARGUMENT;
    SmPartition *p = static_cast<SmPartition *>(handle);
    CUstream s;
    PN2_CU_TRY(driver().GreenCtxStreamCreate(&s, p->ctx[which], CU_STREAM_NON_BLOCKING, 0));
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        g_stream_sms[reinterpret_cast<cudaStream_t>(s)] = p->sms[which];
    }
    *stream = s;
    return PN2_OK;
}

extern "C" int pn2_stream_sm_count(pn2_stream_t stream) { return stream_sm_count(as_stream(stream)); }
