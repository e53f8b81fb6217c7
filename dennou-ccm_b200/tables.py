"""Mapping tables and grids: host mirror of common/grid_mapping_util.f90,
common/grid_mapping_util_jones99.f90 and tool/gmapgen/gmapgen_main.f90 over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib as L


class Grid:
    """1-D axes of a lon-lat grid as gmapgen hands them to the generators (radians, S->N)."""

    def __init__(self, im, jm, x_Lon, y_Lat, x_LonWt, y_LatWt):
        self.im, self.jm = int(im), int(jm)
        self.x_Lon, self.y_Lat, self.x_LonWt, self.y_LatWt = x_Lon, y_Lat, x_LonWt, y_LatWt

    @property
    def n(self):
        return self.im * self.jm


def _grid(fn, im, jm):
    a = [np.zeros(im), np.zeros(jm), np.zeros(im), np.zeros(jm)]
    L.check(fn(im, jm, *[L.dp(x) for x in a]))
    return Grid(im, jm, *a)


def get_LonLatGrid(iMax, jMax):
    """Gaussian grid (ref tool/gmapgen/gmapgen_main.f90:256-307; SPML stand-in)."""
    return _grid(L.lib().dccm_grid_gauss, iMax, jMax)


def regular_LonLatGrid(iMax, jMax):
    """Regular lat-lon grid (cell centres), for the synthetic ocean grids of BASELINE.md."""
    return _grid(L.lib().dccm_grid_regular, iMax, jMax)


def generate_surface_exchange_grid(atm, ocn):
    """ref tool/gmapgen/gmapgen_main.f90:336-405"""
    n = atm.jm + ocn.jm
    lat, wt = np.zeros(n), np.zeros(n)
    jms = C.c_int(0)
    L.check(L.lib().dccm_grid_exchange(atm.jm, L.dp(atm.y_Lat), L.dp(atm.y_LatWt), ocn.jm, L.dp(ocn.y_LatWt),
                                       C.byref(jms), L.dp(lat), L.dp(wt)))
    j = jms.value
    return Grid(atm.im, j, atm.x_Lon.copy(), lat[:j].copy(), atm.x_LonWt.copy(), wt[:j].copy())


class MappingTable:
    """Owns a dccm_table (entries iD jD iS jS coef, 1-based, table-file order)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib().dccm_table_free(self._h)
            self._h = None

    @property
    def n(self):
        return int(L.lib().dccm_table_size(self._h))

    def entries(self):
        n = self.n
        a = [np.zeros(n, np.int32) for _ in range(4)] + [np.zeros(n)]
        L.check(L.lib().dccm_table_get(self._h, *[L.ip(x) for x in a[:4]], L.dp(a[4])))
        return a

    def index(self, GNXS, GNXR):
        """set_mappingTable_interpCoef index arithmetic (ref grid_mapping_util_jones99.f90:498-500)."""
        n = self.n
        send, recv, coef = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        L.check(L.lib().dccm_table_index(self._h, GNXS, GNXR, L.ip(send), L.ip(recv), L.dp(coef)))
        return send, recv, coef

    def write(self, filename, binary=False):
        fn = L.lib().dccm_table_write_bin if binary else L.lib().dccm_table_write_text
        L.check(fn(self._h, str(filename).encode()))

    @staticmethod
    def read(filename, binary=False):
        h = C.c_void_p()
        fn = L.lib().dccm_table_read_bin if binary else L.lib().dccm_table_read_text
        L.check(fn(str(filename).encode(), C.byref(h)))
        return MappingTable(h)


def gen_table_jones99(src, dst, accuracy_order=1, lon_mode=0, rows=None):
    """rows = (j0, j1): only destination rows j0 <= j < j1 (0-based) -- a rank's latitude band."""
    h = C.c_void_p()
    j0, j1 = (0, dst.jm) if rows is None else rows
    L.check(L.lib().dccm_table_gen_jones99_rows(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                L.dp(src.y_LatWt), L.dp(dst.y_LatWt),
                                                accuracy_order, lon_mode, j0 + 1, j1, C.byref(h)))
    return MappingTable(h)


def gen_table_bilinear(src, dst, lon_mode=0, rows=None):
    h = C.c_void_p()
    j0, j1 = (0, dst.jm) if rows is None else rows
    L.check(L.lib().dccm_table_gen_bilinear_rows(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                 dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                 lon_mode, j0 + 1, j1, C.byref(h)))
    return MappingTable(h)


def make_mapping_table(nx_r, ny_r, nx_s, ny_s):
    """ref common/cal_mappingtable.f90:10-49 -- regular grids in degrees, given by their sizes; receiver <- sender."""
    h = C.c_void_p()
    L.check(L.lib().dccm_table_gen_make_mapping_table(nx_r, ny_r, nx_s, ny_s, C.byref(h)))
    return MappingTable(h)


def gen_table_separable(src, dst, conservative, accuracy_order=1, lon_mode=1):
    """The table multiplied out from its separable factors (different longitudes, kind 2) or from its per-row stencil
    (equal longitudes / nx == 1, kind 1), as the kernels do: host check of the forms dccm_remap_create_* build."""
    h = C.c_void_p()
    if conservative:
        L.check(L.lib().dccm_table_gen_jones99_separable(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                         dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                         L.dp(src.y_LatWt), L.dp(dst.y_LatWt),
                                                         accuracy_order, lon_mode, C.byref(h)))
    else:
        L.check(L.lib().dccm_table_gen_bilinear_separable(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                          dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                          lon_mode, C.byref(h)))
    return MappingTable(h)
