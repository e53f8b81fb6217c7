"""Mirror of the ocean / sea-ice glue's element-wise work either side of the remaps
(ref ocn/dccm_ocn_mod.f90:825-851 put side, :978-993 get side), device resident."""
from . import _lib as L


def ocn_put_assemble(SeaSfcTemp, SfcAlbedoAO, SIceCon, SIceSfcTempC, SfcAlbedoAI, IceMaskMin, degC2K,
                     o2s_bil, o2s_cons):
    """torch cuda float64 tensors; writes the five O->S / I->S send layers in place."""
    n, ld = SeaSfcTemp.numel(), o2s_bil.shape[1]
    L.check(L.lib().dccm_ocn_put_assemble_device(
        n, L.tptr(SeaSfcTemp), L.tptr(SfcAlbedoAO), L.tptr(SIceCon), L.tptr(SIceSfcTempC), L.tptr(SfcAlbedoAI),
        float(IceMaskMin), float(degC2K), L.tptr(o2s_bil), L.tptr(o2s_cons), ld, L.current_stream()))


def ocn_get_assemble(o_recv, DensFreshWater, n=None):
    """o_recv (12, ld): returns dict of the six assembled OCN-grid fields."""
    import torch
    ld = o_recv.shape[1]
    n = ld if n is None else n
    names = ("FreshWtFlxS0", "FreshWtFlx0", "WindStressXAI", "WindStressYAI", "SfcHFlxAO0", "DSfcHFlxAODTs")
    out = {k: torch.empty(n, dtype=torch.float64, device=o_recv.device) for k in names}
    L.check(L.lib().dccm_ocn_get_assemble_device(n, L.tptr(o_recv), ld, float(DensFreshWater),
                                                 *[L.tptr(out[k]) for k in names], L.current_stream()))
    return out


class TimeAverage:
    """Jcup RECV_MODE='AVG' of the S->O / S->I variables (ref ocn/dccm_ocn_mod.f90:652-672), device resident:
    put() the surface component's layers at every surface step, get() the mean when the coupling interval
    closes (and start the next interval).  The mean commutes with the (linear) remap, so averaging the 12 packed
    S->O send layers once replaces twelve host-side Jcup buffers."""

    def __init__(self, like):
        import torch
        self.acc = torch.zeros_like(like)
        self.count = 0

    def put(self, x):
        L.check(L.lib().dccm_avg_accumulate_device(L.tptr(self.acc), L.tptr(x), x.numel(), int(self.count == 0),
                                                   L.current_stream()))
        self.count += 1

    def get(self):
        assert self.count > 0, "no put since the last get"
        L.check(L.lib().dccm_avg_finish_device(L.tptr(self.acc), self.acc.numel(), self.count, L.current_stream()))
        self.count = 0
        return self.acc
