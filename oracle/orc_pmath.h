/*
 * orc_pmath.h -- CPU ORACLE (test infrastructure, NOT product code): the "portable" elementary
 * functions of the bulk-flux restatement.
 *
 * WHY.  exp / log / x**y are the only operations of the hot path whose result depends on the
 * run-time library: Intel's libm (the reference's build, sysdep/Makedef.Linux64-intel-impi), glibc
 * and CUDA's libdevice all return faithfully rounded values that differ from one another in the last
 * place, and DSFCM_Util_SfcBulkFlux_Get (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:211-219, :250-259)
 * amplifies one such ulp to a few 1e-12 in the fluxes.  To compare the CUDA path with this oracle
 * BIT FOR BIT the three functions are therefore DEFINED here by one fixed sequence of IEEE-754
 * binary64 additions, multiplications, one division and integer operations -- no fused
 * multiply-add, no table, no library call -- which every IEEE machine evaluates to the same bits.
 * The device code (dennou-ccm_b200/csrc/dccm_pmath.cuh) is a second, independently written text of
 * the SAME sequence; tests compare the two bitwise over dense operand sweeps
 * (tests/test_pmath.py, tests/test_gpu_parity.py).
 *
 * The definition (spec; `*`, `+`, `-`, `/` are round-to-nearest binary64 operations evaluated in
 * the written association):
 *
 *   pm_exp(x):  NaN -> NaN; x > 709.782712893384 -> +Inf; x < -745.1332191019412 -> +0
 *       k  = trunc(x * INVLN2 + (x < 0 ? -0.5 : 0.5))                     (int32)
 *       hi = x - k*LN2HI        (LN2HI has 21 trailing zero bits: the product is exact)
 *       lo = k*LN2LO ;  r = hi - lo ;  z = r*r
 *       q  = E(z) + r*O(z),  E / O = even / odd part of  sum_{n=2..14} r^(n-2)/n!, each a Horner chain in z with
 *            separate multiply and add (coefficients = the doubles nearest 1/n!)
 *       y  = 1 + (hi + (z*q - lo))
 *       result = (y * 2^(k/2)) * 2^(k - k/2)      (k/2 truncated toward zero; the first product is exact, the second
 *                                                  rounds once, into the subnormals when it has to)
 *
 *   pm_log(x):  NaN, x < 0 -> NaN; +-0 -> -Inf; +Inf -> +Inf; subnormal x is first scaled by 2^54
 *       x = m * 2^k with m in [sqrt(2)/2, sqrt(2))  (split at high word 0x3fe6a09e)
 *       f = m - 1 ;  s = f / (2 + f) ;  z = s*s ;  w = z*z
 *       R = z * (E(w) + z*O(w)),  E / O = even / odd part of  sum_{n=1..11} 2 z^(n-1)/(2n+1)   (= log((1+s)/(1-s))/s - 2)
 *       h = (0.5*f)*f
 *       result = k*LN2HI - ((h - (s*(h + R) + k*LN2LO)) - f)
 *
 *   pm_pow(x, y) = y == 0.25 ? sqrt(sqrt(x)) : pm_exp(y * pm_log(x))
 *       (the path raises to GasRDry/CpDry ~ 0.2857 with |y ln x| < 0.1, where this is within 1 ulp
 *        of the correctly rounded power, and to 0.25 -- ref atm/dccm_atm_mod.f90:831)
 *
 * Measured distance from the correctly rounded result (mpmath, tests/test_pmath.py): exp < 0.85 ulp,
 * log < 0.80 ulp over dense sweeps; from glibc's exp / log / pow: <= 1 ulp.
 */
#ifndef ORC_PMATH_H
#define ORC_PMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t orc_pm_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double orc_pm_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

#define ORC_PM_LN2HI  6.93147180369123816490e-01   /* 0x3fe62e42fee00000 */
#define ORC_PM_LN2LO  1.90821492927058770002e-10   /* 0x3dea39ef35793c76 */
#define ORC_PM_INVLN2 1.44269504088896338700e+00   /* 0x3ff71547652b82fe */

static inline double orc_pm_pow2(int e) { return orc_pm_from_bits((uint64_t)(e + 1023) << 52); }   /* e in [-1022, 1023] */

static inline double orc_pm_exp(double x)
{
    if (x != x) return x + x;
    if (x > 709.782712893383973096) return INFINITY;
    if (x < -745.13321910194110842) return 0.0;
    const int k = (int)(x * ORC_PM_INVLN2 + (x < 0.0 ? -0.5 : 0.5));
    const double kd = (double)k;
    const double hi = x - kd * ORC_PM_LN2HI;
    const double lo = kd * ORC_PM_LN2LO;
    const double r = hi - lo;
    const double z = r * r;
    double e = 1.0 / 87178291200.0;              /* even powers of r: 1/14! z^6 + 1/12! z^5 + ... + 1/2! */
    e = e * z + 1.0 / 479001600.0;
    e = e * z + 1.0 / 3628800.0;
    e = e * z + 1.0 / 40320.0;
    e = e * z + 1.0 / 720.0;
    e = e * z + 1.0 / 24.0;
    e = e * z + 0.5;
    double o = 1.0 / 6227020800.0;               /* odd powers of r: 1/13! z^5 + 1/11! z^4 + ... + 1/3! */
    o = o * z + 1.0 / 39916800.0;
    o = o * z + 1.0 / 362880.0;
    o = o * z + 1.0 / 5040.0;
    o = o * z + 1.0 / 120.0;
    o = o * z + 1.0 / 6.0;
    const double q = e + r * o;
    const double y = 1.0 + (hi + (z * q - lo));
    const int k1 = k / 2;                        /* C division truncates toward zero */
    return (y * orc_pm_pow2(k1)) * orc_pm_pow2(k - k1);
}

static inline double orc_pm_log(double x)
{
    uint64_t u = orc_pm_bits(x);
    int k = 0;
    if (x != x) return x + x;
    if (x == 0.0) return -INFINITY;
    if (x < 0.0) return NAN;
    if (x == INFINITY) return x;
    if (u < 0x0010000000000000ull) { x = x * 0x1p54; u = orc_pm_bits(x); k = -54; }
    int hx = (int)(u >> 32);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    if (hx >= 0x6a09e) { k += 1; hx |= 0x3fe00000; }   /* m in [sqrt(2)/2, 1) */
    else hx |= 0x3ff00000;                             /* m in [1, sqrt(2)) */
    const double m = orc_pm_from_bits(((uint64_t)(uint32_t)hx << 32) | (u & 0xffffffffull));
    const double kd = (double)k;
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    double e = 2.0 / 23.0;                       /* even powers of z: 2/23 w^5 + 2/19 w^4 + ... + 2/3 */
    e = e * w + 2.0 / 19.0;
    e = e * w + 2.0 / 15.0;
    e = e * w + 2.0 / 11.0;
    e = e * w + 2.0 / 7.0;
    e = e * w + 2.0 / 3.0;
    double o = 2.0 / 21.0;                       /* odd powers of z: 2/21 w^4 + 2/17 w^3 + ... + 2/5 */
    o = o * w + 2.0 / 17.0;
    o = o * w + 2.0 / 13.0;
    o = o * w + 2.0 / 9.0;
    o = o * w + 2.0 / 5.0;
    const double R = z * (e + z * o);
    const double h = (0.5 * f) * f;
    return kd * ORC_PM_LN2HI - ((h - (s * (h + R) + kd * ORC_PM_LN2LO)) - f);
}

static inline double orc_pm_pow(double x, double y)
{
    if (y == 0.25) return sqrt(sqrt(x));
    return orc_pm_exp(y * orc_pm_log(x));
}

#endif
