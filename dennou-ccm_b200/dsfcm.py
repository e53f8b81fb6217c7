"""Mirror of the DSFCM surface-model modules on the hot path:
DSFCM_Admin_Grid_mod, DSFCM_Admin_Variable_mod, DSFCM_Util_SfcBulkFlux_mod
(ref sfc/DSFCM_Admin_Grid_mod.f90, sfc/DSFCM_Admin_Variable_mod.f90,
sfc/DSFCM_Util_SfcBulkFlux_mod.f90).

Arrays are numpy with REVERSED axis order ((n,) JA, IA) so that the memory layout equals the
Fortran (IA, JA[, n]) column-major arrays the C ABI expects."""
import ctypes as C

import numpy as np

from . import _lib as L

SFC_PROP_MAX = 3   # ref sfc/DSFCM_Admin_Variable_mod.f90:27

OUT3 = ["WindStressX", "WindStressY", "SenHFlx", "QVapMFlx", "LatHFlx",
        "SfcVelTransCoef", "SfcTempTransCoef", "SfcQVapTransCoef",
        "SUwRFlx", "LUwRFlx", "SfcHFlx_ns", "SfcHFlx_sr", "DSfcHFlxDTs"]
IN2 = ["WindU", "WindV", "SfcAirTemp", "QVap1", "SDwRFlx", "LDwRFlx", "SIceCon", "SfcHeight", "SfcPress"]


class DSFCM_Admin_Grid:
    """ref sfc/DSFCM_Admin_Grid_mod.f90:35-52 (1-based bounds, halo 1)."""

    def __init__(self, GNX, GNY):
        self.IM, self.JM = GNX, GNY
        self.IHALO = self.JHALO = 1
        self.IS = self.IHALO + 1
        self.IE = self.IS + self.IM - 1
        self.IA = self.IM + 2 * self.IHALO
        self.JS = self.JHALO + 1
        self.JE = self.JS + self.JM - 1
        self.JA = self.JM + 2 * self.JHALO

    def interior(self):
        return (slice(self.JS - 1, self.JE), slice(self.IS - 1, self.IE))


class DSFCM_Admin_Variable:
    """ref sfc/DSFCM_Admin_Variable_mod.f90:57-80 -- the SFC state arrays (IA,JA,SFC_PROP_MAX)."""

    def __init__(self, grid, fill=0.0):
        IA, JA = grid.IA, grid.JA
        for k in OUT3 + ["SfcTemp", "SfcAlbedo"]:
            setattr(self, "xya_" + k, np.full((SFC_PROP_MAX, JA, IA), fill))
        self.xya_DelVarImplCPL = np.full((4, JA, IA), fill)
        for k in ["SIceCon", "RainFall", "SnowFall", "SDwRFlx", "LDwRFlx"]:
            setattr(self, "xy_" + k, np.full((JA, IA), fill))


def DSFCM_Util_SfcBulkFlux_Get(IA, JA,
                               xya_WindStressX, xya_WindStressY, xya_SenHFlx, xya_QVapMFlx, xya_LatHFlx,
                               xya_SfcVelTransCoef, xya_SfcTempTransCoef, xya_SfcQVapTransCoef,
                               xya_DelVarImplCPL, xya_SUwRFlx, xya_LUwRFlx,
                               xya_SfcHFlx_ns, xya_SfcHFlx_sr, xya_DSfcHFlxDTs,
                               xy_WindU, xy_WindV, xy_SfcAirTemp, xy_QVap1, xy_SDwRFlx, xy_LDwRFlx,
                               xya_ImplCplCoef1, xya_ImplCplCoef2, xya_SfcTemp, xya_SfcAlbedo, xy_SIceCon,
                               a_Sig1Info, xy_SfcHeight, xy_SfcPress):
    """ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:108-151 -- same dummy order (IA, JA come from
    DSFCM_Admin_Grid_mod in the reference).  Host arrays in, host arrays out (in place)."""
    outs = [xya_WindStressX, xya_WindStressY, xya_SenHFlx, xya_QVapMFlx, xya_LatHFlx,
            xya_SfcVelTransCoef, xya_SfcTempTransCoef, xya_SfcQVapTransCoef, xya_DelVarImplCPL,
            xya_SUwRFlx, xya_LUwRFlx, xya_SfcHFlx_ns, xya_SfcHFlx_sr, xya_DSfcHFlxDTs]
    ins = [L.f64(a) for a in (xy_WindU, xy_WindV, xy_SfcAirTemp, xy_QVap1, xy_SDwRFlx, xy_LDwRFlx,
                              xya_ImplCplCoef1, xya_ImplCplCoef2)]
    tail = [L.f64(a) for a in (xy_SIceCon, np.asarray(a_Sig1Info, dtype=np.float64), xy_SfcHeight, xy_SfcPress)]
    for a in outs + [xya_SfcTemp, xya_SfcAlbedo]:
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    L.check(L.lib().dccm_bulkflux_get_host(IA, JA, *[L.dp(a) for a in outs], *[L.dp(a) for a in ins],
                                           L.dp(xya_SfcTemp), L.dp(xya_SfcAlbedo), *[L.dp(a) for a in tail]))


def bulkflux_device(nx, ny, fields, sig1, ld=None, off=0, slot_stride=None):
    """Device-resident form (dccm_bulkflux_device). fields: dict name -> torch cuda float64 tensor
    (missing outputs are not stored)."""
    f = L.SfcFields()
    for n in L.SfcFields._names:
        t = fields.get(n)
        setattr(f, n, None if t is None else t.data_ptr())
    ld = nx if ld is None else ld
    ss = nx * ny if slot_stride is None else slot_stride
    L.check(L.lib().dccm_bulkflux_device(nx, ny, ld, off, ss, C.byref(f), float(sig1), L.current_stream()))
