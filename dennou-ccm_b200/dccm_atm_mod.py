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


def atm_legacy_get_assemble(o2a_recv, cycle_sec, Grav, CpDry, Press0, Press1, TempB1, n=None):
    """Legacy 2-component get side (ref atm/mod_atm.f90:740-775, atm/dcpam_main_mod.f90:1003-1031): o2a_recv (4, ld) =
    remapped SfcTemp**4, SfcAlbedo, SfcEngyFlxMod, SfcSnow; TempB1 (level 1 of the temperature) is corrected in place.
    Returns dict SurfTemp, SurfAlbedo, SurfSnow."""
    import torch
    ld = o2a_recv.shape[1]
    n = ld if n is None else n
    out = {k: torch.empty(n, dtype=torch.float64, device=o2a_recv.device) for k in ("SurfTemp", "SurfAlbedo", "SurfSnow")}
    L.check(L.lib().dccm_atm_legacy_get_assemble_device(
        n, L.tptr(o2a_recv), ld, float(cycle_sec), float(Grav), float(CpDry), L.tptr(Press0), L.tptr(Press1),
        L.tptr(out["SurfTemp"]), L.tptr(out["SurfAlbedo"]), L.tptr(out["SurfSnow"]), L.tptr(TempB1), L.current_stream()))
    return out
