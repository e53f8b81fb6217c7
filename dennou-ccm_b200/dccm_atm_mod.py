"""Mirror of the atmosphere glue's element-wise work on the get side of the S->A remaps
(ref atm/dccm_atm_mod.f90:817-853), device resident."""
from . import _lib as L

STB = 5.670373e-8          # the surface component's Stefan-Boltzmann constant (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:37),
                           # which the glue imports (ref atm/dccm_atm_mod.f90:789)


def atm_get_assemble(a_recv, n=None, StB=STB):
    """a_recv (9, ld) torch cuda float64, the remapped S->A layers: returns dict with xy_SfcTemp = (LUwRFlx/StB)**0.25
    (:831) and the copies the AGCM receives (SfcAlbedo :827, SurfHeatFlux :825, SurfH2OVapFlux :826)."""
    import torch
    ld = a_recv.shape[1]
    n = ld if n is None else n
    names = ("SfcTemp", "SfcAlbedo", "SurfHeatFlux", "SurfH2OVapFlux")
    out = {k: torch.empty(n, dtype=torch.float64, device=a_recv.device) for k in names}
    L.check(L.lib().dccm_atm_get_assemble_device(n, L.tptr(a_recv), ld, float(StB), *[L.tptr(out[k]) for k in names],
                                                 L.current_stream()))
    return out
