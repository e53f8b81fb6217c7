"""ctypes binding of libdccm_b200.so (the C ABI declared in include/dccm_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the
first symbol, and every compute entry point fails with the library's own error message when no
sm_100 device is usable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdccm_b200.so")
CSRC = os.path.join(_HERE, "csrc")

f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p


class DccmError(RuntimeError):
    pass


class SfcFields(C.Structure):
    """dccm_sfc_fields of include/dccm_b200.h (device pointers)."""
    _names = ["WindStressX", "WindStressY", "SenHFlx", "QVapMFlx", "LatHFlx",
              "SfcVelTransCoef", "SfcTempTransCoef", "SfcQVapTransCoef", "DelVarImplCPL",
              "SUwRFlx", "LUwRFlx", "SfcHFlx_ns", "SfcHFlx_sr", "DSfcHFlxDTs",
              "WindU", "WindV", "SfcAirTemp", "QVap1", "SDwRFlx", "LDwRFlx",
              "ImplCplCoef1", "ImplCplCoef2", "SfcTemp", "SfcAlbedo", "SIceCon", "SfcHeight", "SfcPress"]
    _fields_ = [(n, vp) for n in _names]


class AtmSfcFlx(C.Structure):
    """dccm_atm_sfcflx of include/dccm_b200.h (device pointers, 27 in + 11 out)."""
    _in = ["SurfMomFluxX", "SurfMomFluxY", "SurfVelTransCoef", "SurfTempTransCoef", "SurfQVapTransCoef",
           "SurfHumidCoef", "DUDt1", "DVDt1", "DTempDtVDiff1", "DQVapDt1", "HeatFlux0", "QVapFlux0", "ExnerR0",
           "ExnerZ1", "TempN1", "DSurfTempDt", "SnowFrac", "DQVapSatDTempOnLiq", "DQVapSatDTempOnSol",
           "RadLDwFlux0", "RadLUwFlux0", "RadSDwFlux0", "RadSUwFlux0", "DelRadLDwFlux00", "DelRadLDwFlux01",
           "DelRadLUwFlux00", "DelRadLUwFlux01"]
    _out = ["TauXAtm", "TauYAtm", "SensAtm", "LatentAtm", "LDWRFlxAtm", "LUWRFlxAtm", "SDWRFlxAtm", "SUWRFlxAtm",
            "SurfAirTemp", "DSurfLatentFlxDTs", "DSurfHFlxDTs"]
    _fields_ = [(n, vp) for n in _in + _out]


class SrcSeg(C.Structure):
    """dccm_src_seg: a send buffer whose boundary rows live in the neighbouring ranks' buffers."""
    _fields_ = [("lo", vp), ("own", vp), ("hi", vp), ("b0", C.c_int64), ("b1", C.c_int64)]


def build(force=False, verbose=False):
    """Compile libdccm_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs.append(os.path.join(_HERE, "..", "include", "dccm_b200.h"))
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    r = subprocess.run(["make", "-C", CSRC, "-B"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise DccmError("building libdccm_b200.so failed")
    return LIB_PATH


DRIVER_PATH = os.path.join(_HERE, "..", "examples", "exchange_driver")


def build_driver():
    """examples/exchange_driver: one exchange through the C ABI from compiled host code (plain g++)."""
    build()
    r = subprocess.run(["make", "-C", CSRC, "driver"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise DccmError("building examples/exchange_driver failed")
    return os.path.abspath(DRIVER_PATH)


GMAPGEN_PATH = os.path.join(_HERE, "..", "tool", "gmapgen", "gmapgen_main")


def build_gmapgen():
    """tool/gmapgen/gmapgen_main: the reference's table-file generator program over the C ABI (plain g++)."""
    build()
    r = subprocess.run(["make", "-C", CSRC, "gmapgen"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise DccmError("building tool/gmapgen/gmapgen_main failed")
    return os.path.abspath(GMAPGEN_PATH)


_SIGS = {
    "dccm_last_error": (C.c_char_p, []),
    "dccm_build_info": (C.c_char_p, []),
    "dccm_init": (C.c_int, [C.c_int]),
    "dccm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dccm_sync": (C.c_int, [vp]),
    "dccm_host_register": (C.c_int, [vp, C.c_int64]),
    "dccm_host_unregister": (C.c_int, [vp]),
    "dccm_grid_gauss": (C.c_int, [C.c_int, C.c_int, f64p, f64p, f64p, f64p]),
    "dccm_grid_regular": (C.c_int, [C.c_int, C.c_int, f64p, f64p, f64p, f64p]),
    "dccm_grid_exchange": (C.c_int, [C.c_int, f64p, f64p, C.c_int, f64p, C.POINTER(C.c_int), f64p, f64p]),
    "dccm_table_gen_jones99": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                         f64p, f64p, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_bilinear": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                          C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_jones99_rows": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                              f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_bilinear_rows": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                               C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_jones99_separable": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                                   f64p, f64p, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_bilinear_separable": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                                    C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_make_mapping_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_write_text": (C.c_int, [vp, C.c_char_p]),
    "dccm_table_read_text": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "dccm_table_write_bin": (C.c_int, [vp, C.c_char_p]),
    "dccm_table_read_bin": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "dccm_table_size": (C.c_int64, [vp]),
    "dccm_table_get": (C.c_int, [vp, i32p, i32p, i32p, i32p, f64p]),
    "dccm_table_index": (C.c_int, [vp, C.c_int, C.c_int, i32p, i32p, f64p]),
    "dccm_table_free": (None, [vp]),
    "dccm_remap_create": (C.c_int, [C.c_int64, i32p, i32p, f64p, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_remap_create_lonlat": (C.c_int, [C.c_int64, i32p, i32p, f64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(vp)]),
    "dccm_remap_create_jones99": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                            f64p, f64p, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_remap_create_bilinear": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                             C.c_int, C.POINTER(vp)]),
    "dccm_remap_create_jones99_band": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                                 f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(vp)]),
    "dccm_remap_create_bilinear_band": (C.c_int, [C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "dccm_table_gen_band_expanded": (C.c_int, [C.c_int, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p, C.c_int, f64p,
                                               f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(vp)]),
    "dccm_remap_classify": (C.c_int, [C.c_int64, i32p, i32p, f64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "dccm_remap_destroy": (None, [vp]),
    "dccm_remap_nnz": (C.c_int64, [vp]),
    "dccm_remap_kind": (C.c_int, [vp]),
    "dccm_remap_apply_host": (C.c_int, [vp, f64p, C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.c_int]),
    "dccm_remap_apply_device": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
    "dccm_remap_apply_seg_device": (C.c_int, [vp, C.POINTER(SrcSeg), C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
    "dccm_interp_register": (C.c_int, [C.c_int, C.c_int, C.c_int, vp]),
    "dccm_interpolate_data": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f64p,
                                        C.c_int, C.c_int, f64p, C.c_int]),
    "dccm_interp_set_model_name": (C.c_int, [C.c_int, C.c_char_p]),
    "dccm_interpolate_data_named": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, f64p,
                                              C.c_int, C.c_int, f64p, C.c_int]),
    "dccm_f77_set_error_handler": (None, [vp]),
    "dccm_bulkflux_get_host": (C.c_int, [C.c_int, C.c_int] + [f64p] * 28),
    "dccm_bulkflux_device": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                       C.POINTER(SfcFields), C.c_double, vp]),
    "dccm_sfc_exchange_device": (C.c_int, [vp] * 4 + [vp] * 4 + [C.c_int, C.c_double, vp, vp, C.c_int64,
                                           C.POINTER(SfcFields), vp]),
    "dccm_sfc_exchange_seg_device": (C.c_int, [vp] * 4 + [C.POINTER(SrcSeg)] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_double,
                                               vp, vp, C.c_int64, C.POINTER(SfcFields), vp]),
    "dccm_sfc_exchange_rows_device": (C.c_int, [vp] * 4 + [C.POINTER(SrcSeg)] * 4 + [C.c_int64, C.c_int64, C.c_int, C.c_double,
                                                vp, vp, C.c_int64, C.POINTER(SfcFields), C.c_int, C.c_int, vp]),
    "dccm_selftest_pmath_device": (C.c_int, [C.c_int, vp, C.c_int64, C.c_double, vp]),
    "dccm_selftest_fast_arith_device": (C.c_int, [vp, vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dccm_sfc_exchange_config": (C.c_int, [vp, C.c_int, C.c_int]),
    "dccm_sfc_exchange_last_form": (C.c_int, [vp]),
    "dccm_ocn_put_assemble_device": (C.c_int, [C.c_int64] + [vp] * 5 + [C.c_double, C.c_double, vp, vp, C.c_int64, vp]),
    "dccm_ocn_get_assemble_device": (C.c_int, [C.c_int64, vp, C.c_int64, C.c_double] + [vp] * 6 + [vp]),
    "dccm_avg_accumulate_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "dccm_avg_finish_device": (C.c_int, [vp, C.c_int64, C.c_int, vp]),
    "dccm_atm_get_assemble_device": (C.c_int, [C.c_int64, vp, C.c_int64, C.c_double, vp, vp, vp, vp, vp]),
    "dccm_atm_legacy_get_assemble_device": (C.c_int, [C.c_int64, vp, C.c_int64, C.c_double, C.c_double, C.c_double,
                                                      vp, vp, vp, vp, vp, vp, vp]),
    "dccm_atm_store_surf_flx_device": (C.c_int, [C.c_int64, C.POINTER(AtmSfcFlx), C.c_double, C.c_double, C.c_double, vp]),
    "dccm_vdiff_create": (C.c_int, [C.c_int] * 5 + [C.c_double] * 4 + [C.POINTER(vp)]),
    "dccm_vdiff_destroy": (None, [vp]),
    "dccm_vdiff_set_mode": (C.c_int, [vp, C.c_int]),
    "dccm_vdiff_set_coef_stride": (C.c_int, [vp, C.c_int64]),
    "dccm_vdiff_forward_host": (C.c_int, [vp] + [f64p] * 18),
    "dccm_vdiff_backward_host": (C.c_int, [vp] + [f64p] * 4),
    "dccm_vdiff_forward_device": (C.c_int, [vp] + [vp] * 18 + [vp]),
    "dccm_vdiff_backward_device": (C.c_int, [vp] + [vp] * 4 + [vp, vp]),
    "dccm_vdiff_redo_total": (C.c_int, [vp, C.POINTER(C.c_int64)]),
    "dccm_vdiff_forward_cols_device": (C.c_int, [vp] + [vp] * 18 + [C.c_int64, C.c_int64, vp]),
    "dccm_vdiff_backward_cols_device": (C.c_int, [vp] + [vp] * 4 + [vp, C.c_int64, C.c_int64, vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DccmError(f"{LIB_PATH} is missing: run __graft_entry__.build() "
                            "(there is no CPU fallback for the exchange path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


F77_INTERPOLATE_DATA = ("interpolate_data_", None, [C.c_char_p, C.c_char_p] + [i32p] * 3 + [f64p] + [i32p] * 2 + [f64p]
                        + [i32p] * 3 + [C.c_size_t, C.c_size_t])


def f77_interpolate_data():
    """the bare Fortran external `interpolate_data_` the library exports for Jcup (every argument by reference,
    CHARACTER lengths appended by value)"""
    fn = getattr(lib(), F77_INTERPOLATE_DATA[0])
    fn.restype, fn.argtypes = F77_INTERPOLATE_DATA[1], F77_INTERPOLATE_DATA[2]
    return fn


def declared_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc != 0:
        raise DccmError(f"libdccm_b200 error {rc}: {lib().dccm_last_error().decode()}")


def dp(a):
    """double* of a C-contiguous float64 numpy array."""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(f64p)


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], "need contiguous int32"
    return a.ctypes.data_as(i32p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def tptr(t):
    """device pointer of a torch CUDA float64 tensor (or None)."""
    if t is None:
        return None
    import torch
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need contiguous cuda float64"
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
