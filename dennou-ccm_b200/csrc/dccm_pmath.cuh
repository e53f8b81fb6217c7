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
// Distance from the correctly rounded result: exp <= 0.85 ulp, log <= 0.78 ulp (measured against mpmath,
// tests/test_pmath.py) -- inside the 1 ulp any libm promises.
//
//   pexp(x):  k = trunc(x*INVLN2 -+ 0.5); hi = x - k*LN2HI (exact); lo = k*LN2LO; r = hi - lo; z = r*r
//             q = E(z) + r*O(z),  E / O = the even / odd part of  sum_{n=2..14} r^(n-2)/n!  (two Horner chains in z,
//             multiply and add rounded separately -- half the dependent depth of one chain in r)
//             y = 1 + (hi + (z*q - lo));   result = (y * 2^(k/2)) * 2^(k - k/2)   (k/2 truncated: the first product is
//             exact, the second rounds once -- also into the subnormals)
//             NaN -> NaN; x > 709.782712893384 -> +Inf; x < -745.1332191019412 -> +0
//   plog(x):  x = m*2^k, m in [sqrt(2)/2, sqrt(2)) (subnormal x scaled by 2^54 first); f = m - 1; s = f/(2+f); z = s*s; w = z*z
//             R = z*(E(w) + z*O(w)),  E / O = the even / odd part of  sum_{n=1..11} 2 z^(n-1)/(2n+1);  h = (0.5*f)*f
//             result = k*LN2HI - ((h - (s*(h+R) + k*LN2LO)) - f)
//             NaN, x < 0 -> NaN; +-0 -> -Inf; +Inf -> +Inf
//   ppow(x,y) = exp(y*log(x));   x**0.25 = sqrt(sqrt(x))
//
// Every product and sum is written with __dmul_rn / __dadd_rn / __dsub_rn so no compiler flag can contract them, and
// the special operands are handled by selects at the end, not by branches: a call is one basic block that the
// scheduler can interleave with its neighbours (the surface kernel is bound by the latency of dependent fp64 chains).
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
    const bool nan = x != x, big = x > 709.782712893383973096, tiny = x < -745.13321910194110842;
    const double xc = (nan || big || tiny) ? 0.0 : x;             // the main path always runs on an ordinary operand
    const int k = __double2int_rz(add(mul(xc, INVLN2), xc < 0.0 ? -0.5 : 0.5));
    const double kd = (double)k;
    const double hi = sub(xc, mul(kd, LN2HI));
    const double lo = mul(kd, LN2LO);
    const double r = sub(hi, lo);
    const double z = mul(r, r);
    const double *c = kExpC;                                      // c[i] = 1/(14-i)!: c[12] is the r^0 term, c[0] the r^12 term
    double e = c[0], o = c[1];                                    // even powers r^12, r^10, ..., r^0; odd powers r^11, ..., r^1
#pragma unroll
    for (int i = 2; i < 13; i += 2) e = add(mul(e, z), c[i]);
#pragma unroll
    for (int i = 3; i < 13; i += 2) o = add(mul(o, z), c[i]);
    const double q = add(e, mul(r, o));
    const double y = add(1.0, add(hi, sub(mul(z, q), lo)));
    const int k1 = k / 2;
    const double res = mul(mul(y, pow2(k1)), pow2(k - k1));
    return nan ? add(x, x) : (big ? __longlong_as_double(0x7ff0000000000000LL) : (tiny ? 0.0 : res));
}

// `Arith` supplies the one division (FastArith: branch-free, with the caller's redo when it was not acceptable).
template <class Arith>
__device__ __forceinline__ double plog(double x, Arith &ar)
{
    using namespace pm;
    const bool nan = x != x, zero = x == 0.0, neg = x < 0.0;
    const bool inf = __double2hiint(x) >= 0x7ff00000 && !nan && !neg;
    const bool special = nan || zero || neg || inf;
    const double x1 = special ? 1.0 : x;                          // the main path always runs on a positive finite operand
    const bool sub_ = __double2hiint(x1) < 0x00100000;            // subnormal: scale by 2^54
    const double xs = sub_ ? mul(x1, 18014398509481984.0) : x1;
    int hx = __double2hiint(xs);
    int k = (sub_ ? -54 : 0) + (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const bool up = hx >= 0x6a09e;                                // m in [sqrt(2)/2, 1) : [1, sqrt(2))
    k += up ? 1 : 0;
    hx |= up ? 0x3fe00000 : 0x3ff00000;
    const double m = __hiloint2double(hx, __double2loint(xs));
    const double kd = (double)k;
    const double f = sub(m, 1.0);
    const double s = ar.div(f, add(2.0, f));
    const double z = mul(s, s);
    const double w = mul(z, z);
    const double *c = kLogC;                                      // c[i] = 2/(23-2i): c[10] is the z^0 term, c[0] the z^10 term
    double e = c[0], o = c[1];                                    // even powers z^10, ..., z^0; odd powers z^9, ..., z^1
#pragma unroll
    for (int i = 2; i < 11; i += 2) e = add(mul(e, w), c[i]);
#pragma unroll
    for (int i = 3; i < 11; i += 2) o = add(mul(o, w), c[i]);
    const double R = mul(z, add(e, mul(z, o)));
    const double h = mul(mul(0.5, f), f);
    const double res = sub(mul(kd, LN2HI), sub(sub(h, add(mul(s, add(h, R)), mul(kd, LN2LO))), f));
    return nan ? add(x, x)
               : (zero ? __longlong_as_double(0xfff0000000000000LL)
                       : (neg ? __longlong_as_double(0x7ff8000000000000LL) : (inf ? x : res)));
}

template <class Arith>
__device__ __forceinline__ double ppow(double x, double y, Arith &ar)
{
    return pexp(pm::mul(y, plog(x, ar)));
}

// x**0.25 (ref atm/dccm_atm_mod.f90:831, atm/mod_atm.f90:743): two correctly rounded square roots
__device__ __forceinline__ double pfourth_root(double x) { return sqrt(sqrt(x)); }

}  // namespace dccm
