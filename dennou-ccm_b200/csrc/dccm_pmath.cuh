// dccm_pmath.cuh -- exp / log / x**y of the surface-flux column as FIXED sequences of IEEE-754 binary64
// operations.
//
// exp, log and `**` (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:211-219, :250-259; atm/dccm_atm_mod.f90:831) are the
// only operations of the exchange whose bits depend on a run-time library: Intel libm, glibc and libdevice
// each return a faithfully rounded value, not the same one, and the stability functions amplify that last
// place to a few 1e-12 in the fluxes.  The library therefore does not call libdevice for them: the three
// functions are defined by the sequence below (round-to-nearest `+`, `*`, one `/`, integer exponent
// manipulation; no fused multiply-add, no table), which has one answer on every IEEE machine, and a host
// model that wants the exchange reproducible across CPU and GPU builds evaluates the same sequence
// (DESIGN.md section 5 gives it as a specification; the test suite carries its own C text of it).
// Distance from the correctly rounded result: exp <= 0.81 ulp, log <= 0.78 ulp (measured against mpmath,
// tests/test_pmath.py) -- inside the 1 ulp any libm promises.
//
//   pexp(x):  k = trunc(x*INVLN2 -+ 0.5); hi = x - k*LN2HI (exact); lo = k*LN2LO; r = hi - lo
//             q = sum_{n=2..14} r^(n-2)/n!  (Horner, multiply and add rounded separately)
//             y = 1 + (hi + ((r*r)*q - lo));   result = y * 2^k
//   plog(x):  x = m*2^k, m in [sqrt(2)/2, sqrt(2)); f = m - 1; s = f/(2+f); z = s*s
//             R = z * sum_{n=1..11} 2 z^(n-1)/(2n+1); h = (0.5*f)*f
//             result = k*LN2HI - ((h - (s*(h+R) + k*LN2LO)) - f)
//   ppow(x,y) = exp(y*log(x));   x**0.25 = sqrt(sqrt(x))
//
// Every product and sum is written with __dmul_rn / __dadd_rn / __dsub_rn so no compiler flag can contract them.
#pragma once

namespace dccm {

namespace pm {
// Polynomial coefficients live in constant memory: the compiler then fetches two of them per LDCU.128 instead of
// materialising each 64-bit literal with two UMOV instructions (the surface kernel is instruction-issue bound).
// 1/n!, n = 14 .. 2 and 2/(2n+1), n = 11 .. 1: the divisions are folded by the compiler to the nearest double.
static __constant__ double kExpC[13] = {1.0 / 87178291200.0, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0,
                                        1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
                                        1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
static __constant__ double kLogC[11] = {2.0 / 23.0, 2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0,
                                        2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0};
constexpr double LN2HI = 6.93147180369123816490e-01;    // upper 32 bits of ln 2: k*LN2HI is exact for |k| < 2^21
constexpr double LN2LO = 1.90821492927058770002e-10;    // ln 2 - LN2HI
constexpr double INVLN2 = 1.44269504088896338700e+00;

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double pow2(int e) { return __hiloint2double((e + 1023) << 20, 0); }   // 2^e, e in [-1022, 1023]
}  // namespace pm

__device__ __forceinline__ double pexp(double x)
{
    using namespace pm;
    if (x != x) return add(x, x);
    if (x > 709.782712893383973096) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -745.13321910194110842) return 0.0;
    const int k = __double2int_rz(add(mul(x, INVLN2), x < 0.0 ? -0.5 : 0.5));
    const double kd = (double)k;
    const double hi = sub(x, mul(kd, LN2HI));
    const double lo = mul(kd, LN2LO);
    const double r = sub(hi, lo);
    const double *c = kExpC;
    double q = c[0];
#pragma unroll
    for (int i = 1; i < 13; i++) q = add(mul(q, r), c[i]);
    const double y = add(1.0, add(hi, sub(mul(mul(r, r), q), lo)));
    if (k > 1023) return mul(mul(y, pow2(1023)), pow2(k - 1023));
    if (k < -1021) return mul(mul(y, pow2(k + 1000)), pow2(-1000));
    return __hiloint2double(__double2hiint(y) + (k << 20), __double2loint(y));
}

// `Arith` supplies the one division (FastArith: branch-free, with the caller's redo when it was not acceptable).
template <class Arith>
__device__ __forceinline__ double plog(double x, Arith &ar)
{
    using namespace pm;
    if (x != x) return add(x, x);
    if (x == 0.0) return __longlong_as_double(0xfff0000000000000LL);
    if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
    int hx = __double2hiint(x), k = 0;
    if (hx >= 0x7ff00000) return x;                                   // +Inf
    if (hx < 0x00100000) { x = mul(x, 18014398509481984.0); hx = __double2hiint(x); k = -54; }   // subnormal: * 2^54
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    if (hx >= 0x6a09e) { k += 1; hx |= 0x3fe00000; }                 // m in [sqrt(2)/2, 1)
    else hx |= 0x3ff00000;                                            // m in [1, sqrt(2))
    const double m = __hiloint2double(hx, __double2loint(x));
    const double kd = (double)k;
    const double f = sub(m, 1.0);
    const double s = ar.div(f, add(2.0, f));
    const double z = mul(s, s);
    const double *c = kLogC;
    double p = c[0];
#pragma unroll
    for (int i = 1; i < 11; i++) p = add(mul(p, z), c[i]);
    const double R = mul(z, p);
    const double h = mul(mul(0.5, f), f);
    return sub(mul(kd, LN2HI), sub(sub(h, add(mul(s, add(h, R)), mul(kd, LN2LO))), f));
}

template <class Arith>
__device__ __forceinline__ double ppow(double x, double y, Arith &ar)
{
    return pexp(pm::mul(y, plog(x, ar)));
}

// x**0.25 (ref atm/dccm_atm_mod.f90:831, atm/mod_atm.f90:743): two correctly rounded square roots
__device__ __forceinline__ double pfourth_root(double x) { return sqrt(sqrt(x)); }

}  // namespace dccm
