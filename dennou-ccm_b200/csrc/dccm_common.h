// dccm_common.h -- shared helpers of libdccm_b200 (error reporting, CUDA checks).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dccm_b200.h"

namespace dccm {

void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#ifdef __CUDACC__
}  // namespace dccm
#include <nvtx3/nvToolsExt.h>
namespace dccm {
// NVTX range around every stage of the exchange (SURVEY.md section 5 "Tracing"): shows up in Nsight Systems / ncu
// timelines under the entry point's name; header-only NVTX v3, a no-op when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define DCCM_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return dccm::fail(DCCM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                              cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)
#endif

// Grow-only device scratch buffer used by the *_host entry points.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

int ensure_device();   // lazily dccm_init(current device); fails loudly without a GPU
int num_sms();

}  // namespace dccm
