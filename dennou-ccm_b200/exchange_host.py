"""Exchange step with HOST inputs and outputs, slab-pipelined over PCIe.

The resident step (exchange.SurfaceExchange) assumes the coupled fields live in HBM.  When the
models still run on the host (the reference situation), every exchange has to move the forward
solve's inputs host->device (19.8 GB for T1279L26) and the tendencies + remapped fields back
(7.3 GB); the kernels are ~2 % of that.  This class hides the device->host leg behind the
host->device leg: the grids are cut into latitude slabs (the same BandPlan that shards across
GPUs, used here for "virtual ranks" on ONE device), and slab s runs

    H2D(s) | forward(s) | surface kernel(s-1) | remaps + backward(s-2) | D2H(s-2)

on three streams, so results of the first slabs stream back while the inputs of the later slabs
are still arriving.  Halo rows between slabs are plain device copies.  Results are bit-identical
to the resident path (checked by bench.py on every run).

The host side is slab-contiguous, which is how the reference holds it: the atmosphere is
decomposed into latitude bands over MPI ranks (ref atm/dccm_atm_mod.f90:172-176, :953-1001).
"""
from . import sharding
from .dcpam_sfc_implicit_coupling_mod import IN_ORDER

ATM_SFC = ("WindU", "WindV", "SfcAirTemp", "QVap1", "SfcPress", "LDwRFlx", "SDwRFlx", "RainFall", "SnowFall")
OCN_SFC = ("SfcTempO", "SfcTempI", "SIceCon", "SfcAlbedoO", "SfcAlbedoI")


class HostPipelinedExchange:
    def __init__(self, A, O, S, kmax, ncmax=1, index_h2ovap=1, nslab=12, device=None, fast=False):
        import torch
        self.torch = torch
        self.n = nslab
        self.plan = sharding.BandPlan(A, O, S, nslab)
        self.slabs = [sharding.ShardedExchange(A, O, S, kmax, ncmax, index_h2ovap, rank=s, world=nslab,
                                               plan=self.plan, dist=None, device=device, fast=fast)
                      for s in range(nslab)]
        self.dev = self.slabs[0].dev
        self.K, self.nc = kmax, ncmax
        self.h2d, self.d2h = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        z = lambda *shape: torch.empty(shape, dtype=torch.float64, device=self.dev)
        pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True)
        self.d_in, self.h_in, self.h_out = [], [], []
        for s, ex in enumerate(self.slabs):
            nA, nO = ex.A.n, ex.O.n
            shapes = {k: ((ncmax, kmax + 1, nA) if k == "QMixFlux" else
                          (kmax, nA) if k in ("zExner", "Height") else (kmax + 1, nA)) for k in IN_ORDER}
            shapes.update({"a:" + k: (nA,) for k in ATM_SFC})
            shapes.update({"o:" + k: (nO,) for k in OCN_SFC})
            self.d_in.append({k: z(*sh) for k, sh in shapes.items()})
            self.h_in.append({k: pin(*sh) for k, sh in shapes.items()})
            outs = {"DUDt": (kmax, nA), "DVDt": (kmax, nA), "DTempDt": (kmax, nA), "DQMixDt": (ncmax, kmax, nA),
                    "a_recv": (9, nA), "o_recv": (12, nO)}
            self.h_out.append({k: pin(*sh) for k, sh in outs.items()})
            ex.col_in = {k: self.d_in[s][k] for k in IN_ORDER}
        # halo rows between neighbouring slabs: (dst buffer, c0, c1, src buffer, source cell offset)
        self.pull_in, self.pull_out = [[] for _ in range(nslab)], [[] for _ in range(nslab)]
        for s, ex in enumerate(self.slabs):
            for names, g, lst in ((("a2s_bil", "a2s_cons"), "A", self.pull_in), (("o2s_bil", "o2s_cons"), "O", self.pull_in),
                                  (("s2a", "s2o"), "S", self.pull_out)):
                im = self.plan.grid[g].im
                for peer, kind, c0, c1 in self.plan.halo_messages(g, s):
                    if kind != "recv":
                        continue
                    shift = (self.plan.ext[g][s][0] - self.plan.ext[g][peer][0]) * im
                    for name in names:
                        lst[s].append((getattr(ex, name), c0, c1, getattr(self.slabs[peer], name), shift))
        self.h2d_bytes = sum(t.numel() * 8 for d in self.h_in for t in d.values())
        self.d2h_bytes = sum(t.numel() * 8 for d in self.h_out for t in d.values())

    def bands(self, s):
        return self.plan.bands["A"][s], self.plan.bands["O"][s]

    @staticmethod
    def _pull(items):
        for dst, c0, c1, src, shift in items:
            dst[:, c0:c1] = src[:, c0 + shift:c1 + shift]

    def step(self):
        """one exchange: host inputs (self.h_in) -> host outputs (self.h_out); returns after enqueueing,
        synchronise the device (or self.d2h) before reading h_out"""
        torch, n = self.torch, self.n
        cs = torch.cuda.current_stream(self.dev)
        self.h2d.wait_stream(cs)              # previous exchange finished reading the device inputs
        cs.wait_stream(self.d2h)              # ... and its results have left the device
        ev_in = []
        with torch.cuda.stream(self.h2d):
            for s in range(n):
                for k, h in self.h_in[s].items():
                    self.d_in[s][k].copy_(h, non_blocking=True)
                e = torch.cuda.Event()
                e.record(self.h2d)
                ev_in.append(e)
        for t in range(n + 2):
            if t < n:
                ex = self.slabs[t]
                cs.wait_event(ev_in[t])
                d = self.d_in[t]
                ex.set_inputs(ex.col_in, {k: d["a:" + k] for k in ATM_SFC}, {k: d["o:" + k] for k in OCN_SFC})
                ex.forward()
            if 0 <= t - 1 < n:
                self._pull(self.pull_in[t - 1])
                self.slabs[t - 1].sfc_fused()
            if 0 <= t - 2 < n:
                s = t - 2
                ex = self.slabs[s]
                self._pull(self.pull_out[s])
                ex.remap_from_sfc()
                ex.backward()
                e = torch.cuda.Event()
                e.record(cs)
                self.d2h.wait_event(e)
                with torch.cuda.stream(self.d2h):
                    o = self.h_out[s]
                    for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"):
                        o[k].copy_(ex.tend[k], non_blocking=True)
                    o["a_recv"].copy_(ex.a_recv, non_blocking=True)
                    o["o_recv"].copy_(ex.o_recv, non_blocking=True)

    def synchronize(self):
        self.d2h.synchronize()
        self.torch.cuda.current_stream(self.dev).synchronize()
