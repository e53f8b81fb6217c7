// dccm_runtime.cu -- device selection, error reporting, scratch buffers.
#include <cuda_runtime.h>

#include "dccm_common.h"

namespace dccm {

static thread_local std::string g_err;
static int g_device = -1;
static int g_sms = 0;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return DCCM_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    DCCM_CUDA_TRY(cudaMalloc(&p, bytes));
    cap = bytes;
    return DCCM_OK;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

int ensure_device()
{
    if (g_device >= 0) return DCCM_OK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(DCCM_ERR_CUDA, "no usable CUDA device (%s); libdccm_b200 has no CPU fallback",
                    cudaGetErrorString(e));
    return dccm_init(dev);
}

int num_sms() { return g_sms > 0 ? g_sms : 148; }

}  // namespace dccm

using namespace dccm;

extern "C" const char *dccm_last_error(void)
{
    return dccm::g_err.c_str();
}

extern "C" const char *dccm_build_info(void)
{
    return "libdccm_b200 sm_100a fp64 (CUDA " __DATE__ ")";
}

extern "C" int dccm_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(DCCM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return DCCM_OK;
}

extern "C" int dccm_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(DCCM_ERR_CUDA, "no CUDA device visible (%s); libdccm_b200 has no CPU fallback",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(DCCM_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    DCCM_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    DCCM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DCCM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    g_device = device;
    g_sms = prop.multiProcessorCount;
    return DCCM_OK;
}

extern "C" int dccm_sync(void *stream)
{
    DCCM_CUDA_TRY(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    return DCCM_OK;
}

// Page-lock a caller-owned host array (cudaHostRegister) so that the *_host entry points can overlap their
// H2D / kernel / D2H chunks; pageable arrays work too, the copies then serialise through the driver's staging.
extern "C" int dccm_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return fail(DCCM_ERR_ARG, "dccm_host_register: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    DCCM_CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return DCCM_OK;
}

extern "C" int dccm_host_unregister(void *ptr)
{
    if (!ptr) return fail(DCCM_ERR_ARG, "dccm_host_unregister: null pointer");
    DCCM_CUDA_TRY(cudaHostUnregister(ptr));
    return DCCM_OK;
}
