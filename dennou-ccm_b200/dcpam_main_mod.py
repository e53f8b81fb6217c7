"""Mirror of the atmosphere component's surface-flux bookkeeping after the backward solve
(dcpam_StoreAtmSurfFlxInfo, ref atm/dcpam_main_mod.f90:1040-1114), device resident."""
from . import _lib as L


def dcpam_StoreAtmSurfFlxInfo(fields, LatentHeat, CpDry, DelTime):
    """fields: dict name -> torch cuda float64 tensor of n columns holding the 27 inputs named in
    include/dccm_b200.h (dccm_atm_sfcflx); returns a dict with the 11 outputs (xy_TauXAtm ... xy_DSurfHFlxDTs)."""
    import torch
    f = L.AtmSfcFlx()
    first = fields[L.AtmSfcFlx._in[0]]
    n = first.numel()
    for k in L.AtmSfcFlx._in:
        assert fields[k].numel() == n, k
        setattr(f, k, fields[k].data_ptr())
        assert fields[k].is_cuda and fields[k].dtype == torch.float64 and fields[k].is_contiguous(), k
    out = {k: torch.empty(n, dtype=torch.float64, device=first.device) for k in L.AtmSfcFlx._out}
    for k, t in out.items():
        setattr(f, k, t.data_ptr())
    import ctypes as C
    L.check(L.lib().dccm_atm_store_surf_flx_device(n, C.byref(f), float(LatentHeat), float(CpDry), float(DelTime),
                                                   L.current_stream()))
    return out
