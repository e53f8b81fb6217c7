"""Mirror of module interpolation_data_latlon_mod + the external interpolate_data subroutine
(ref common/interpolation_data_latlon_mod.f90, common/interpolate_data.f90).

Jcup is external: where the reference asks Jcup for the local operation indices
(jcup_get_local_operation_index, :140-146) and the local coefficients (jcup_set_local_coef /
jcup_recv_coef, :240-267) the caller passes them in.  On one rank local == global."""
import ctypes as C

import numpy as np

from . import _lib as L

_model_ids = {}          # component name -> Jcup component number (1-based)
_pending = {}            # (recv, send, tag) -> (send_index, recv_index, n_send, n_recv)
_handles = {}            # (recv, send, tag) -> RemapOperator


class RemapOperator:
    """One operation_index_type entry (ref :67-83) living on the device as row-sorted CSR."""

    def __init__(self, send_index, recv_index, coef, n_send, n_recv, gnxs=0, gnxr=0):
        """gnxs / gnxr: longitudes per row of the source / destination grid (0 = unknown): lets the
        library store a zonally repeating table as one stencil per latitude row (kind 1)."""
        send_index, recv_index, coef = L.i32(send_index), L.i32(recv_index), L.f64(coef)
        assert len(send_index) == len(recv_index) == len(coef)
        self.n_send, self.n_recv = int(n_send), int(n_recv)
        h = C.c_void_p()
        L.check(L.lib().dccm_remap_create_lonlat(len(coef), L.ip(send_index), L.ip(recv_index), L.dp(coef),
                                                 self.n_send, self.n_recv, int(gnxs), int(gnxr), C.byref(h)))
        self._h = h

    @classmethod
    def from_grids(cls, src, dst, conservative, accuracy_order=1, lon_mode=1, rows=None, src_rows=None):
        """Generate-and-create in one step (dccm_remap_create_jones99 / _bilinear): the table is never built.
        Different longitudes give the separable form (kind 2); equal longitudes a zonal stencil (kind 1).
        rows = (j0, j1), src_rows = (e0, e1): the operator of a latitude band -- destination rows [j0, j1) only, source
        cells numbered inside a buffer that holds source rows [e0, e1) (dccm_remap_create_*_band)."""
        self = cls.__new__(cls)
        self.n_send, self.n_recv = src.n, dst.n
        h = C.c_void_p()
        if rows is not None:
            (j0, j1), (e0, e1) = rows, src_rows
            self.n_send, self.n_recv = (e1 - e0) * src.im, (j1 - j0) * dst.im
            if conservative:
                L.check(L.lib().dccm_remap_create_jones99_band(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                               dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                               L.dp(src.y_LatWt), L.dp(dst.y_LatWt),
                                                               accuracy_order, lon_mode, j0, j1, e0, e1 - e0, C.byref(h)))
            else:
                L.check(L.lib().dccm_remap_create_bilinear_band(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                                dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                                lon_mode, j0, j1, e0, e1 - e0, C.byref(h)))
        elif conservative:
            L.check(L.lib().dccm_remap_create_jones99(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                      dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                      L.dp(src.y_LatWt), L.dp(dst.y_LatWt),
                                                      accuracy_order, lon_mode, C.byref(h)))
        else:
            L.check(L.lib().dccm_remap_create_bilinear(src.im, L.dp(src.x_Lon), src.jm, L.dp(src.y_Lat),
                                                       dst.im, L.dp(dst.x_Lon), dst.jm, L.dp(dst.y_Lat),
                                                       lon_mode, C.byref(h)))
        self._h = h
        return self

    @property
    def kind(self):
        return int(L.lib().dccm_remap_kind(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib().dccm_remap_destroy(self._h)
            self._h = None

    @property
    def nnz(self):
        return int(L.lib().dccm_remap_nnz(self._h))

    def apply_host(self, send_data, rn1=None, rn2=None, num_of_data=None):
        """send_data: numpy (sn2, sn1) C-order == Fortran send_data(sn1, sn2); returns (rn2, rn1)."""
        send_data = L.f64(send_data)
        sn2, sn1 = send_data.shape
        rn1 = self.n_recv if rn1 is None else rn1
        rn2 = sn2 if rn2 is None else rn2
        nd = sn2 if num_of_data is None else num_of_data
        recv = np.full((rn2, rn1), np.nan)
        L.check(L.lib().dccm_remap_apply_host(self._h, L.dp(send_data), sn1, sn2, L.dp(recv), rn1, rn2, nd))
        return recv

    def apply(self, send, recv=None, num_of_data=None):
        """Device-resident: send torch (sn2, sn1) cuda float64 -> recv (rn2, rn1) on the current stream."""
        import torch
        sn2, sn1 = send.shape
        if recv is None:
            recv = torch.empty((sn2, self.n_recv), dtype=torch.float64, device=send.device)
        rn2, rn1 = recv.shape
        nd = min(sn2, rn2) if num_of_data is None else num_of_data
        L.check(L.lib().dccm_remap_apply_device(self._h, L.tptr(send), sn1, L.tptr(recv), rn1, rn2, nd,
                                                L.current_stream()))
        return recv


def interpolation_data_latlon_Init(num_of_model, num_of_mapping_tag, my_model_id, model_names=None):
    """ref :94-104. model_names maps component names to Jcup component numbers
    (jcup_get_comp_num_from_name, :289); default ATM=1, OCN=2, SFC=3 (:51-53)."""
    _model_ids.clear(); _pending.clear(); _handles.clear()
    _model_ids.update(model_names or {"ATM": 1, "OCN": 2, "SFC": 3})


def _key(recv_model, send_model, tag):
    try:
        return (_model_ids[recv_model], _model_ids[send_model], int(tag))
    except KeyError as e:   # jcup_error aborts in the reference
        raise L.DccmError(f"unknown component name {e}") from None


def set_operation_index(recv_model_name, send_model_name, mapping_tag=1, *,
                        send_data_index, recv_data_index, num_of_send_grid, num_of_recv_grid):
    """ref :116-154 (the keyword arguments are what jcup_get_local_operation_index returns)."""
    _pending[_key(recv_model_name, send_model_name, mapping_tag)] = (
        L.i32(send_data_index), L.i32(recv_data_index), int(num_of_send_grid), int(num_of_recv_grid))


def set_interpolate_coef(send_comp_name, recv_comp_name, coef_owncomp_name, mapping_tag, coefS_global=None):
    """ref :219-270. Builds the device operator once indices and coefficients are both known."""
    k = _key(recv_comp_name, send_comp_name, mapping_tag)
    if coefS_global is None:
        raise L.DccmError(f"set_interpolate_coef: NO coefS. send={send_comp_name}, recv={recv_comp_name}")
    if k not in _pending:
        raise L.DccmError("set_interpolate_coef called before set_operation_index")
    s, r, ns, nr = _pending[k]
    op = RemapOperator(s, r, coefS_global, ns, nr)
    _handles[k] = op
    L.check(L.lib().dccm_interp_register(k[0], k[1], k[2], op._h))
    return op


def interpolate_data_latlon(recv_model, send_model, send_data, recv_data, num_of_data, grid_num, exchange_tag=None):
    """ref :274-306. send_data (sn2, sn1), recv_data (rn2, rn1) numpy C-order; recv_data is overwritten."""
    k = _key(recv_model, send_model, grid_num)
    send_data = L.f64(send_data)
    assert recv_data.dtype == np.float64 and recv_data.flags["C_CONTIGUOUS"]
    sn2, sn1 = send_data.shape
    rn2, rn1 = recv_data.shape
    L.check(L.lib().dccm_interpolate_data(k[0], k[1], k[2], sn1, sn2, L.dp(send_data), rn1, rn2,
                                          L.dp(recv_data), num_of_data))


def interpolate_data(recv_model, send_model, mapping_tag, sn1, sn2, send_data,
                     rn1, rn2, recv_data, num_of_data, tn=0, exchange_tag=None):
    """The symbol Jcup calls back (ref common/interpolate_data.f90:1-17), same argument order."""
    assert send_data.shape == (sn2, sn1) and recv_data.shape == (rn2, rn1)
    interpolate_data_latlon(recv_model, send_model, send_data, recv_data, num_of_data, mapping_tag, exchange_tag)
