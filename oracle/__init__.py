"""ctypes binding of the CPU oracle (oracle/dccm_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/dccm_oracle.h).  May be imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and
only as the checker / the timed CPU baseline; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libdccm_oracle.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("dccm_oracle.c", "dccm_oracle.h", "orc_pmath.h", "Makefile")]
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= max(os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.orc_table_new.restype = C.c_void_p
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_table_n.restype = C.c_int64
        L.orc_table_n.argtypes = [C.c_void_p]
        L.orc_table_copy.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p, _f64p]
        L.orc_gauss_grid.argtypes = [C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.orc_regular_grid.argtypes = [C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.orc_gen_jones99.argtypes = [C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p,
                                      _f64p, _f64p, C.c_int, C.c_int, C.c_void_p]
        L.orc_gen_jones99_rows.argtypes = [C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p,
                                           _f64p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_gen_bilinear_rows.argtypes = [C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_gen_bilinear.argtypes = [C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f64p,
                                       C.c_int, C.c_void_p]
        L.orc_make_mapping_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_exchange_grid.argtypes = [C.c_int, _f64p, _f64p, C.c_int, _f64p,
                                        C.POINTER(C.c_int), _f64p, _f64p]
        L.orc_table_write_text.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_table_read_text.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_remap_apply.argtypes = [C.c_int64, _i32p, _i32p, _f64p, _f64p, C.c_int, C.c_int,
                                      _f64p, C.c_int, C.c_int, C.c_int]
        L.orc_bulkflux.argtypes = [C.c_int, C.c_int] + [_f64p] * 28
        L.orc_set_math.argtypes = [C.c_int]
        L.orc_pm_exp_v.argtypes = [C.c_int64, _f64p, _f64p]
        L.orc_pm_log_v.argtypes = [C.c_int64, _f64p, _f64p]
        L.orc_pm_pow_v.argtypes = [C.c_int64, _f64p, C.c_double, _f64p]
        L.orc_vdiff_new.restype = C.c_void_p
        L.orc_vdiff_new.argtypes = [C.c_int] * 5 + [C.c_double] * 4
        L.orc_vdiff_free.argtypes = [C.c_void_p]
        L.orc_vdiff_forward.argtypes = [C.c_void_p] + [_f64p] * 18
        L.orc_vdiff_backward.argtypes = [C.c_void_p] + [_f64p] * 4
        L.orc_vdiff_get_diag.argtypes = [C.c_void_p, C.c_int, _f64p]
        L.orc_ocn_put_assemble.argtypes = [C.c_int64] + [_f64p] * 5 + [C.c_double, C.c_double, _f64p, _f64p]
        L.orc_ocn_get_assemble.argtypes = [C.c_int64] + [_f64p] * 8 + [C.c_double] + [_f64p] * 6
        L.orc_avg_accumulate.argtypes = [_f64p, _f64p, C.c_int64, C.c_int]
        L.orc_avg_finish.argtypes = [_f64p, C.c_int64, C.c_int]
        L.orc_atm_store_surf_flx.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.c_double, C.c_double, C.c_double]
        L.orc_atm_sfc_temp.argtypes = [C.c_int64, _f64p, C.c_double, _f64p]
        L.orc_atm_legacy_get.argtypes = [C.c_int64, _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
                                         _f64p, _f64p, _f64p, _f64p, _f64p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


class Grid:
    """1-D axes of a lon-lat grid in the form gmapgen hands to the generators (radians)."""

    def __init__(self, im, jm, x_Lon, y_Lat, x_LonWt, y_LatWt):
        self.im, self.jm = im, jm
        self.x_Lon, self.y_Lat, self.x_LonWt, self.y_LatWt = x_Lon, y_Lat, x_LonWt, y_LatWt

    @property
    def n(self):
        return self.im * self.jm


def gauss_grid(im, jm):
    a = [np.zeros(im), np.zeros(jm), np.zeros(im), np.zeros(jm)]
    lib().orc_gauss_grid(im, jm, *a)
    return Grid(im, jm, *a)


def regular_grid(im, jm):
    a = [np.zeros(im), np.zeros(jm), np.zeros(im), np.zeros(jm)]
    lib().orc_regular_grid(im, jm, *a)
    return Grid(im, jm, *a)


def exchange_grid(atm, ocn):
    """ref tool/gmapgen/gmapgen_main.f90:336-405"""
    n = atm.jm + ocn.jm
    lat, wt = np.zeros(n), np.zeros(n)
    jms = C.c_int(0)
    lib().orc_exchange_grid(atm.jm, atm.y_Lat, atm.y_LatWt, ocn.jm, ocn.y_LatWt, C.byref(jms), lat, wt)
    j = jms.value
    return Grid(atm.im, j, atm.x_Lon.copy(), lat[:j].copy(), atm.x_LonWt.copy(), wt[:j].copy())


class Table:
    def __init__(self, iD, jD, iS, jS, coef):
        self.iD, self.jD, self.iS, self.jS, self.coef = iD, jD, iS, jS, coef

    @property
    def n(self):
        return len(self.coef)

    def to_index(self, gnxs, gnxr):
        """ref common/grid_mapping_util_jones99.f90:498-500 (1-based linear indices)"""
        recv = (self.iD + gnxr * (self.jD - 1)).astype(np.int32)
        send = (self.iS + gnxs * (self.jS - 1)).astype(np.int32)
        return send, recv, self.coef.copy()


def _take(t):
    n = lib().orc_table_n(t)
    a = [np.zeros(n, np.int32) for _ in range(4)] + [np.zeros(n)]
    lib().orc_table_copy(t, *a)
    lib().orc_table_free(t)
    return Table(*a)


def gen_jones99(src, dst, order=1, lon_mode=0, rows=None):
    """rows = (j0, j1), 0-based half-open destination rows: only those lines of the full table"""
    t = lib().orc_table_new()
    j0, j1 = rows if rows is not None else (0, dst.jm)
    rc = lib().orc_gen_jones99_rows(src.im, src.x_Lon, src.jm, src.y_Lat, dst.im, dst.x_Lon, dst.jm, dst.y_Lat,
                                    src.y_LatWt, dst.y_LatWt, order, lon_mode, j0 + 1, j1, t)
    if rc != 0:
        lib().orc_table_free(t)
        raise RuntimeError(f"orc_gen_jones99 failed rc={rc}")
    return _take(t)


def gen_bilinear(src, dst, lon_mode=0, rows=None):
    t = lib().orc_table_new()
    j0, j1 = rows if rows is not None else (0, dst.jm)
    rc = lib().orc_gen_bilinear_rows(src.im, src.x_Lon, src.jm, src.y_Lat, dst.im, dst.x_Lon, dst.jm, dst.y_Lat,
                                     lon_mode, j0 + 1, j1, t)
    if rc != 0:
        lib().orc_table_free(t)
        raise RuntimeError(f"orc_gen_bilinear failed rc={rc}")
    return _take(t)


def make_mapping_table(nx_r, ny_r, nx_s, ny_s):
    """ref common/cal_mappingtable.f90:10-49 (regular grids in degrees, entries with coef > 0 only)"""
    t = lib().orc_table_new()
    rc = lib().orc_make_mapping_table(nx_r, ny_r, nx_s, ny_s, t)
    if rc != 0:
        lib().orc_table_free(t)
        raise RuntimeError(f"orc_make_mapping_table failed rc={rc}")
    return _take(t)


def write_table(tab, filename):
    with open(filename, "w") as f:
        for k in range(tab.n):
            f.write("%12d%12d%12d%12d  %24.16E\n" % (tab.iD[k], tab.jD[k], tab.iS[k], tab.jS[k], tab.coef[k]))


def read_table(filename):
    t = lib().orc_table_new()
    rc = lib().orc_table_read_text(filename.encode(), t)
    if rc != 0:
        lib().orc_table_free(t)
        raise FileNotFoundError(filename)
    return _take(t)


def remap_apply(send_index, recv_index, coef, send, rn1, rn2=None, num_of_data=None, recv=None):
    """send: (sn2, sn1) C-order == Fortran send_data(sn1, sn2). Returns recv (rn2, rn1); `recv` = a caller-owned
    array to reuse (timing runs), default a fresh NaN-filled one."""
    send = np.ascontiguousarray(send, dtype=np.float64)
    sn2, sn1 = send.shape
    rn2 = sn2 if rn2 is None else rn2
    nd = sn2 if num_of_data is None else num_of_data
    if recv is None:
        recv = np.full((rn2, rn1), np.nan)
    assert recv.shape == (rn2, rn1) and recv.dtype == np.float64 and recv.flags["C_CONTIGUOUS"]
    lib().orc_remap_apply(len(coef), np.ascontiguousarray(send_index, np.int32),
                          np.ascontiguousarray(recv_index, np.int32),
                          np.ascontiguousarray(coef, np.float64), send, sn1, sn2, recv, rn1, rn2, nd)
    return recv


BULK_OUT3 = ["WindStressX", "WindStressY", "SenHFlx", "QVapMFlx", "LatHFlx",
             "SfcVelTransCoef", "SfcTempTransCoef", "SfcQVapTransCoef"]
BULK_OUT3B = ["SUwRFlx", "LUwRFlx", "SfcHFlx_ns", "SfcHFlx_sr", "DSfcHFlxDTs"]
BULK_IN2 = ["WindU", "WindV", "SfcAirTemp", "QVap1", "SDwRFlx", "LDwRFlx"]


def bulkflux(IA, JA, inp, fill=np.nan, out=None):
    """inp: dict with (JA,IA) arrays WindU.. and (4,JA,IA) ImplCplCoef1/2, (3,JA,IA) SfcTemp/SfcAlbedo
    (slots 1,2 set), SIceCon, SfcHeight, SfcPress (JA,IA), Sig1Info (2,).
    Returns dict of outputs (arrays (3|4,JA,IA)); SfcTemp/SfcAlbedo are updated copies.
    out: the dict of a previous call to write into again (timing runs: the reference's module arrays exist before the call)."""
    if out is None:
        out = {}
        for k in BULK_OUT3 + BULK_OUT3B:
            out[k] = np.full((3, JA, IA), fill)
        out["DelVarImplCPL"] = np.full((4, JA, IA), fill)
        out["SfcTemp"] = np.ascontiguousarray(inp["SfcTemp"], dtype=np.float64).copy()
        out["SfcAlbedo"] = np.ascontiguousarray(inp["SfcAlbedo"], dtype=np.float64).copy()
    else:
        out["SfcTemp"][:2] = inp["SfcTemp"][:2]
        out["SfcAlbedo"][:2] = inp["SfcAlbedo"][:2]
    g = lambda k: np.ascontiguousarray(inp[k], dtype=np.float64)
    lib().orc_bulkflux(IA, JA,
                       out["WindStressX"], out["WindStressY"], out["SenHFlx"], out["QVapMFlx"], out["LatHFlx"],
                       out["SfcVelTransCoef"], out["SfcTempTransCoef"], out["SfcQVapTransCoef"],
                       out["DelVarImplCPL"], out["SUwRFlx"], out["LUwRFlx"],
                       out["SfcHFlx_ns"], out["SfcHFlx_sr"], out["DSfcHFlxDTs"],
                       g("WindU"), g("WindV"), g("SfcAirTemp"), g("QVap1"), g("SDwRFlx"), g("LDwRFlx"),
                       g("ImplCplCoef1"), g("ImplCplCoef2"),
                       out["SfcTemp"], out["SfcAlbedo"], g("SIceCon"),
                       g("Sig1Info"), g("SfcHeight"), g("SfcPress"))
    return out


VDIFF_IN = ["MomFluxX", "MomFluxY", "HeatFlux", "QMixFlux", "Press", "zExner", "rExner",
            "VirTemp", "Height", "VelDiffCoef", "TempDiffCoef", "QMixDiffCoef"]


class VDiff:
    """ref atm/dcpam_sfc_implicit_coupling_mod.f90 (module state = the three matrices)"""

    def __init__(self, imax, jmax, kmax, ncmax, index_h2ovap, Grav, CpDry, GasRDry, DelTime):
        self.shape = (imax, jmax, kmax, ncmax)
        self.h = lib().orc_vdiff_new(imax, jmax, kmax, ncmax, index_h2ovap, Grav, CpDry, GasRDry, DelTime)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vdiff_free(self.h)
            self.h = None

    def forward(self, inp, out=None):
        """out: dict of preallocated result arrays to reuse (timing runs: the reference's arrays exist before the
        call); default = fresh NaN-filled arrays, so that anything left unwritten shows in the tests"""
        imax, jmax, K, nc = self.shape
        ncol = imax * jmax
        g = lambda k: np.ascontiguousarray(inp[k], dtype=np.float64)
        if out is None:
            out = {"DUDt": np.full((K, ncol), np.nan), "DVDt": np.full((K, ncol), np.nan),
                   "DTempDt": np.full((K, ncol), np.nan), "DQMixDt": np.full((nc, K, ncol), np.nan),
                   "ImplCplCoef1": np.full((4, ncol), np.nan), "ImplCplCoef2": np.full((4, ncol), np.nan)}
        lib().orc_vdiff_forward(self.h, *[g(k) for k in VDIFF_IN],
                                out["DUDt"], out["DVDt"], out["DTempDt"], out["DQMixDt"],
                                out["ImplCplCoef1"], out["ImplCplCoef2"])
        return out

    def backward(self, DUDt, DVDt, DTempDt, DQMixDt, inplace=False):
        """inplace: overwrite the (contiguous float64) arguments as the reference does (timing runs)"""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (DUDt, DVDt, DTempDt, DQMixDt)]
        if not inplace:
            a = [x.copy() for x in a]
        lib().orc_vdiff_backward(self.h, *a)
        return a

    def diag(self, which):
        imax, jmax, K, nc = self.shape
        out = np.zeros((K, imax * jmax))
        lib().orc_vdiff_get_diag(self.h, which, out)
        return out


def ocn_put_assemble(SeaSfcTemp, AlbAO, SIceCon, SIceSfcTempC, AlbAI, IceMaskMin, degC2K):
    """ref ocn/dccm_ocn_mod.f90:825-836 -> (SIceSfcTemp [K], SIceAlbedo)"""
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (SeaSfcTemp, AlbAO, SIceCon, SIceSfcTempC, AlbAI)]
    n = a[0].size
    ti, ai = np.full(n, np.nan), np.full(n, np.nan)
    lib().orc_ocn_put_assemble(n, *a, IceMaskMin, degC2K, ti, ai)
    return ti, ai


def ocn_get_assemble(ns, sr, dFdT, Snow, Rain, Evap, WSX, WSY, DensFreshWater):
    """ref ocn/dccm_ocn_mod.f90:978-993"""
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (ns, sr, dFdT, Snow, Rain, Evap, WSX, WSY)]
    n = a[0].size
    names = ("FreshWtFlxS0", "FreshWtFlx0", "WindStressXAI", "WindStressYAI", "SfcHFlxAO0", "DSfcHFlxAODTs")
    out = {k: np.full(n, np.nan) for k in names}
    lib().orc_ocn_get_assemble(n, *a, DensFreshWater, *[out[k] for k in names])
    return out


ATM_SFCFLX_IN = ("SurfMomFluxX", "SurfMomFluxY", "SurfVelTransCoef", "SurfTempTransCoef", "SurfQVapTransCoef",
                 "SurfHumidCoef", "DUDt1", "DVDt1", "DTempDtVDiff1", "DQVapDt1", "HeatFlux0", "QVapFlux0", "ExnerR0",
                 "ExnerZ1", "TempN1", "DSurfTempDt", "SnowFrac", "DQVapSatDTempOnLiq", "DQVapSatDTempOnSol",
                 "RadLDwFlux0", "RadLUwFlux0", "RadSDwFlux0", "RadSUwFlux0", "DelRadLDwFlux00", "DelRadLDwFlux01",
                 "DelRadLUwFlux00", "DelRadLUwFlux01")
ATM_SFCFLX_OUT = ("TauXAtm", "TauYAtm", "SensAtm", "LatentAtm", "LDWRFlxAtm", "LUWRFlxAtm", "SDWRFlxAtm", "SUWRFlxAtm",
                  "SurfAirTemp", "DSurfLatentFlxDTs", "DSurfHFlxDTs")


def atm_store_surf_flx(fields, LatentHeat, CpDry, DelTime):
    """dcpam_StoreAtmSurfFlxInfo, ref atm/dcpam_main_mod.f90:1068-1112: dict of the 27 inputs -> dict of 11 outputs"""
    a = [np.ascontiguousarray(fields[k], dtype=np.float64).ravel() for k in ATM_SFCFLX_IN]
    n = a[0].size
    out = {k: np.full(n, np.nan) for k in ATM_SFCFLX_OUT}
    pin = (C.c_void_p * len(a))(*[x.ctypes.data for x in a])
    pout = (C.c_void_p * len(out))(*[out[k].ctypes.data for k in ATM_SFCFLX_OUT])
    lib().orc_atm_store_surf_flx(n, pin, pout, LatentHeat, CpDry, DelTime)
    return out


def atm_sfc_temp(LUwRFlx, StB=5.670373e-8):
    """ref atm/dccm_atm_mod.f90:831"""
    x = np.ascontiguousarray(LUwRFlx, dtype=np.float64)
    out = np.empty_like(x)
    lib().orc_atm_sfc_temp(x.size, x, StB, out)
    return out


def atm_legacy_get(SfcTemp4, SfcSnow, SfcEngyFlxMod, cycle_sec, Grav, CpDry, Press0, Press1, TempB1):
    """ref atm/mod_atm.f90:743, :772-773; atm/dcpam_main_mod.f90:1026-1028.  Returns (SurfTemp, SurfSnow, TempB1 corrected)."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    n = f(SfcTemp4).size
    st, sn, tb = np.empty(n), np.empty(n), f(TempB1).copy()
    lib().orc_atm_legacy_get(n, f(SfcTemp4), f(SfcSnow), f(SfcEngyFlxMod), cycle_sec, Grav, CpDry, f(Press0), f(Press1), st, sn, tb)
    return st, sn, tb


def time_average(puts):
    """Jcup RECV_MODE='AVG' restated: mean of the fields put during one coupling interval, accumulated in put order"""
    acc = np.full(np.asarray(puts[0]).size, np.nan)
    for i, x in enumerate(puts):
        lib().orc_avg_accumulate(acc, np.ascontiguousarray(x, dtype=np.float64).ravel(), acc.size, int(i == 0))
    lib().orc_avg_finish(acc, acc.size, len(puts))
    return acc.reshape(np.asarray(puts[0]).shape)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def set_math(portable=True):
    """exp / log / x**y of the bulk flux: portable fixed IEEE sequences (orc_pmath.h, default) or libm."""
    lib().orc_set_math(1 if portable else 0)


def get_math():
    return bool(lib().orc_get_math())


def pm_exp(x):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.empty_like(x)
    lib().orc_pm_exp_v(x.size, x.reshape(-1), y.reshape(-1)); return y


def pm_log(x):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.empty_like(x)
    lib().orc_pm_log_v(x.size, x.reshape(-1), y.reshape(-1)); return y


def pm_pow(x, e):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.empty_like(x)
    lib().orc_pm_pow_v(x.size, x.reshape(-1), float(e), y.reshape(-1)); return y
