"""Mirror of module dcpam_sfc_implicit_coupling_mod (ref atm/dcpam_sfc_implicit_coupling_mod.f90).

Arrays are numpy / torch with REVERSED axis order: xyr_* -> (kmax+1, jmax*imax),
xyz_* -> (kmax, jmax*imax), xyzf_/xyrf_ -> (ncmax, levels, jmax*imax), xya_ -> (4, jmax*imax),
i.e. the memory layout of the Fortran (0:imax-1, 1:jmax, level[, n]) arrays."""
import ctypes as C

import numpy as np

from . import _lib as L

IN_ORDER = ["MomFluxX", "MomFluxY", "HeatFlux", "QMixFlux", "Press", "zExner", "rExner",
            "VirTemp", "Height", "VelDiffCoef", "TempDiffCoef", "QMixDiffCoef"]


class SfcImplicitCoupling:
    """dcpam_sfc_implicit_coupling_Init (ref :420-426): the swept matrices persist in the handle
    between Forward and Backward like the module's `save` arrays (:16-18)."""

    def __init__(self, imax, jmax, kmax, ncmax, IndexH2OVap, Grav, CpDry, GasRDry, DelTime, fast=False):
        self.imax, self.jmax, self.kmax, self.ncmax = imax, jmax, kmax, ncmax
        h = C.c_void_p()
        L.check(L.lib().dccm_vdiff_create(imax, jmax, kmax, ncmax, IndexH2OVap, Grav, CpDry, GasRDry, DelTime,
                                          C.byref(h)))
        self._h = h
        if fast:
            self.set_fast(True)

    def set_fast(self, fast):
        L.check(L.lib().dccm_vdiff_set_mode(self._h, 1 if fast else 0))

    def redo_total(self):
        """columns re-solved with plain IEEE operators so far (operands outside the branch-free division's range)"""
        n = C.c_int64(0)
        L.check(L.lib().dccm_vdiff_redo_total(self._h, C.byref(n)))
        return int(n.value)

    def set_coef_stride(self, slot_stride):
        L.check(L.lib().dccm_vdiff_set_coef_stride(self._h, int(slot_stride)))

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib().dccm_vdiff_destroy(self._h)
            self._h = None

    # ---- host (drop-in) ----
    def VDiffForward(self, inp):
        """SfcImplicitCoupling_VDiffForward (ref :72-79). inp: dict of host arrays keyed by IN_ORDER."""
        K, nc, ncol = self.kmax, self.ncmax, self.imax * self.jmax
        a = [L.f64(inp[k]) for k in IN_ORDER]
        out = {"DUDt": np.full((K, ncol), np.nan), "DVDt": np.full((K, ncol), np.nan),
               "DTempDt": np.full((K, ncol), np.nan), "DQMixDt": np.full((nc, K, ncol), np.nan),
               "ImplCplCoef1": np.full((4, ncol), np.nan), "ImplCplCoef2": np.full((4, ncol), np.nan)}
        L.check(L.lib().dccm_vdiff_forward_host(self._h, *[L.dp(x) for x in a],
                                                *[L.dp(out[k]) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt",
                                                                         "ImplCplCoef1", "ImplCplCoef2")]))
        return out

    def VDiffBackward(self, xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt):
        """SfcImplicitCoupling_VDiffBackward (ref :25-27): in place on the four host arrays."""
        for a in (xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        L.check(L.lib().dccm_vdiff_backward_host(self._h, L.dp(xyz_DUDt), L.dp(xyz_DVDt), L.dp(xyz_DTempDt),
                                                 L.dp(xyzf_DQMixDt)))

    # ---- device-resident ----
    def forward_device(self, inp, out):
        """inp/out: dicts of torch cuda float64 tensors (same keys as the host form)."""
        L.check(L.lib().dccm_vdiff_forward_device(
            self._h, *[L.tptr(inp[k]) for k in IN_ORDER],
            *[L.tptr(out[k]) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt", "ImplCplCoef1", "ImplCplCoef2")],
            L.current_stream()))

    def forward_cols_device(self, inp, out, c0, c1):
        """columns [c0, c1) only (latitude-slab pipelining)"""
        L.check(L.lib().dccm_vdiff_forward_cols_device(
            self._h, *[L.tptr(inp[k]) for k in IN_ORDER],
            *[L.tptr(out[k]) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt", "ImplCplCoef1", "ImplCplCoef2")],
            int(c0), int(c1), L.current_stream()))

    def backward_cols_device(self, out, level1, c0, c1):
        L.check(L.lib().dccm_vdiff_backward_cols_device(
            self._h, *[L.tptr(out[k]) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt")],
            L.tptr(level1), int(c0), int(c1), L.current_stream()))

    def backward_device(self, out, level1=None):
        L.check(L.lib().dccm_vdiff_backward_device(
            self._h, *[L.tptr(out[k]) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt")],
            L.tptr(level1), L.current_stream()))
