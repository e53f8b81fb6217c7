"""The portable exp / log / x**y that make the bulk flux reproducible bit for bit (oracle/orc_pmath.h on the
host, dennou-ccm_b200/csrc/dccm_pmath.cuh on the device; specification in DESIGN.md section 5).

CPU-side facts about the DEFINITION (no GPU needed):
  * distance from the correctly rounded result, measured with mpmath at 200 bits: < 1 ulp;
  * distance from glibc's exp / log / pow: <= 1 ulp on every sampled operand;
  * the device header's text, compiled for the host through a shim that maps its CUDA intrinsics to their
    IEEE meaning, returns the oracle's bits on millions of operands (the GPU run of the same comparison is
    tests/test_gpu_parity.py::test_portable_exp_log_pow_device_equals_oracle_bitwise);
  * what switching the oracle between the portable functions and libm does to a whole bulk-flux evaluation.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAPPA = 8.3144621 / 1.8e-2 / 1616.0       # GasRDry / CpDry of the surface component (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:41-51)
SPECIAL = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 1.0, -1.0,
                    709.78, 709.79, -745.0, -745.2, -708.4, -740.0, 1.7976931348623157e308])


def _ulps(a, b):
    return int(np.abs(a.view(np.int64) - b.view(np.int64)).max())


def _ulp_error(f, exact, xs):
    import mpmath as mp
    mp.mp.prec = 200
    worst = 0.0
    for x, v in zip(xs, f(xs)):
        t = exact(mp.mpf(float(x)))
        u = float(np.spacing(abs(float(t))))
        worst = max(worst, float(abs((mp.mpf(float(v)) - t) / mp.mpf(u))))
    return worst


def test_distance_from_the_correctly_rounded_value(orc):
    mp = pytest.importorskip("mpmath")
    rng = np.random.default_rng(5)
    xe = np.concatenate([rng.uniform(-30, 30, 6000), rng.uniform(-1, 1, 6000), rng.uniform(-700, 700, 2000)])
    xl = np.concatenate([rng.uniform(0.5, 2, 6000), np.exp(rng.uniform(-700, 700, 4000)), rng.uniform(0.99, 1.01, 4000)])
    e_exp = _ulp_error(orc.pm_exp, mp.exp, xe)
    e_log = _ulp_error(orc.pm_log, mp.log, xl)
    # the operand ranges of the path: Exner functions of pressures within a factor 3 of 1000 hPa, fourth roots of sigma T^4 / sigma
    e_pow = _ulp_error(lambda a: orc.pm_pow(a, KAPPA), lambda t: t ** mp.mpf(KAPPA), rng.uniform(0.3, 3.0, 6000))
    e_r4 = _ulp_error(lambda a: orc.pm_pow(a, 0.25), lambda t: t ** mp.mpf(0.25), rng.uniform(1e8, 1e11, 6000))
    print(f"ulp error vs correctly rounded: exp {e_exp:.3f} log {e_log:.3f} x**kappa {e_pow:.3f} x**0.25 {e_r4:.3f}")
    assert e_exp < 0.9 and e_log < 0.9 and e_pow < 1.0 and e_r4 < 0.9


def test_distance_from_glibc(orc):
    rng = np.random.default_rng(6)
    n = 400000
    xe = np.concatenate([rng.uniform(-700, 700, n), rng.uniform(-1, 1, n)])
    xl = np.concatenate([np.exp(rng.uniform(-700, 700, n)), rng.uniform(0.5, 2, n)])
    xp = rng.uniform(0.3, 3.0, n)
    d = (_ulps(orc.pm_exp(xe), np.exp(xe)), _ulps(orc.pm_log(xl), np.log(xl)), _ulps(orc.pm_pow(xp, KAPPA), np.power(xp, KAPPA)))
    print("max distance from glibc (ulp): exp %d log %d pow %d" % d)
    assert max(d) <= 1
    with np.errstate(all="ignore"):
        for got, want in ((orc.pm_exp(SPECIAL), np.exp(SPECIAL)), (orc.pm_log(SPECIAL), np.log(SPECIAL))):
            assert np.array_equal(np.isnan(got), np.isnan(want))
            m = ~np.isnan(want)
            assert _ulps(got[m], want[m]) <= 1


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "pmath_shim.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "native", "pmath_host_shim.cpp")], check=True)
    L = C.CDLL(so)
    f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.shim_pmath.argtypes = [C.c_int, f64, C.c_int64, C.c_double, f64]

    def run(which, x, y=0.0):
        out = np.empty_like(x)
        L.shim_pmath(which, x, x.size, y, out)
        return out
    return run


def test_device_text_equals_oracle_text_on_the_host(orc, shim):
    rng = np.random.default_rng(7)
    n = 1 << 20
    xe = np.concatenate([rng.uniform(-750, 720, n), rng.uniform(-1, 1, n), SPECIAL])
    xl = np.concatenate([np.exp(rng.uniform(-745, 709, n)), rng.uniform(0.5, 2, n), SPECIAL])
    xp = np.concatenate([rng.uniform(0.3, 3, n), np.exp(rng.uniform(-30, 30, n)), SPECIAL])
    for which, x, y, ref in ((0, xe, 0.0, orc.pm_exp), (1, xl, 0.0, orc.pm_log),
                             (2, xp, KAPPA, lambda a: orc.pm_pow(a, KAPPA)), (3, xl, 0.0, lambda a: orc.pm_pow(a, 0.25))):
        got, want = shim(which, x, y), ref(x)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), which


def test_bulkflux_portable_vs_libm(orc, dccm):
    """What the choice of exp / log / pow does to DSFCM_Util_SfcBulkFlux_Get: the oracle evaluated with the portable
    functions against the same oracle with glibc's.  The routine is ill-conditioned (Louis stability functions, the
    air-sea potential-temperature difference), so the last place of the elementary functions shows up as ~1e-13 .. 1e-11
    relative to the layer's magnitude -- the distance any two conforming builds of the reference have from each other."""
    import importlib
    from exchange_ref import floor_rel
    from test_oracle_kat import _bulk_inputs
    S = importlib.import_module("dennou-ccm_b200.synthetic")
    g = dccm.tables.get_LonLatGrid(128, 64)
    IA, JA, inp = _bulk_inputs(S, g)
    assert orc.get_math()
    a = orc.bulkflux(IA, JA, inp)
    orc.set_math(False)
    try:
        b = orc.bulkflux(IA, JA, inp)
    finally:
        orc.set_math(True)
    worst = {k: floor_rel(a[k][:, 1:-1, 1:-1].reshape(a[k].shape[0], -1), b[k][:, 1:-1, 1:-1].reshape(b[k].shape[0], -1)) for k in a}
    print("portable vs libm, |d| / max(|x|, 1e-3 max|layer|):", {k: float("%.1e" % v) for k, v in worst.items()})
    assert max(worst.values()) <= 2e-11
    assert any(v > 0.0 for v in worst.values())      # they do differ: that is why the definition is needed
