// Test-only: compiles the DEVICE text of the portable elementary functions (csrc/dccm_pmath.cuh) for the host by
// mapping the handful of CUDA intrinsics it uses to their IEEE meaning, so the CPU suite can compare that text
// with the oracle's own C text (oracle/orc_pmath.h) bit for bit without a GPU.  g++ -O2 -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __constant__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline double __hiloint2double(int hi, int lo)
{
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &u, 8); return x;
}
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline int __double2int_rz(double x) { return (int)x; }
#include "../../dennou-ccm_b200/csrc/dccm_pmath.cuh"

struct HostArith { double div(double a, double b) { volatile double r = a / b; return r; } };

extern "C" void shim_pmath(int which, const double *x, int64_t n, double y, double *out)
{
    HostArith ar;
    for (int64_t i = 0; i < n; i++) {
        if (which == 0) out[i] = dccm::pexp(x[i]);
        else if (which == 1) out[i] = dccm::plog(x[i], ar);
        else if (which == 2) out[i] = dccm::ppow(x[i], y, ar);
        else out[i] = dccm::pfourth_root(x[i]);
    }
}
