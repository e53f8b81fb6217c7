"""One atmosphere <-> surface <-> ocean coupling exchange, device resident (SURVEY.md 8d metric 1):

    K3 forward (ATM)  ->  remap A->S (17 layers) + O/I->S (5)  ->  K2 bulk flux (SFC)
                      ->  remap S->A (9) + S->O/S->I (12)       ->  K4 backward (ATM)

The field lists, their grouping by mapping table and their order follow the reference glue
(ref sfc/dccm_sfc_mod.f90:449-466 get-vars, :764-784 put-vars; atm/dccm_atm_mod.f90:697-712,
:817-835; ocn/dccm_ocn_mod.f90:625-645; Appendix A of SURVEY.md): 43 remapped layers per exchange.
Everything between the forward inputs and the backward outputs stays in HBM; the host-side
pack/unpack/put/get staging of the reference glue disappears.

torch owns the device buffers and the stream; every kernel is the library's own
(libdccm_b200.so) launched through the C ABI on torch's current stream.
"""
import numpy as np

from . import _lib as L
from . import dsfcm, tables
from .dcpam_sfc_implicit_coupling_mod import SfcImplicitCoupling, IN_ORDER
from .interpolation_data_latlon_mod import RemapOperator

# layer lists; mapping tag 1 = bilinear, 2 = conservative
A2S_BIL = ["WindU", "WindV", "SfcAirTemp", "QVap1", "SfcPress"] + ["ImplCplCoef1"] * 4 + ["ImplCplCoef2"] * 4   # 13
A2S_CONS = ["LDwRFlx", "SDwRFlx", "RainFall", "SnowFall"]                                                      # 4
O2S_BIL = ["SfcTempO", "SfcTempI"]                                                                             # 2
O2S_CONS = ["SIceCon", "SfcAlbedoO", "SfcAlbedoI"]                                                             # 3
S2A_CONS = ["LUwRFlx3", "SUwRFlx3", "SenHFlx3", "QVapMFlx3"]                                                   # 4
S2A_BIL = ["SfcAlbedo3"] + ["DelVarImplCPL"] * 4                                                               # 5
S2O_CONS = ["SfcHFlx_ns1", "SfcHFlx_sr1", "SnowFall", "RainFall", "Evap1", "-WindStressX3", "-WindStressY3",
            "SfcHFlx_ns2", "SfcHFlx_sr2", "Evap2"]                                                             # 10
S2O_BIL = ["DSfcHFlxDTs1", "DSfcHFlxDTs2"]                                                                     # 2
N_LAYERS = len(A2S_BIL) + len(A2S_CONS) + len(O2S_BIL) + len(O2S_CONS) + len(S2A_CONS) + len(S2A_BIL) \
    + len(S2O_CONS) + len(S2O_BIL)
assert N_LAYERS == 43


def build_tables(A, O, S, order_as=1, lon_mode=1):
    """The 8 tables gmapgen writes for the 3-component mode (ref tool/gmapgen/gmapgen_main.f90:100-149),
    as (send_index, recv_index, coef) triples keyed by direction and kind."""
    t = {}
    for key, s, d in (("as", A, S), ("sa", S, A), ("os", O, S), ("so", S, O)):
        order = order_as if key == "as" else 1
        t[key + "_cons"] = tables.gen_table_jones99(s, d, order, lon_mode).index(s.im, d.im)
        t[key + "_bil"] = tables.gen_table_bilinear(s, d, lon_mode).index(s.im, d.im)
    return t


class SurfaceExchange:
    """Device-resident exchange step for one rank's share of the grids (whole grids on 1 GPU).

    members > 1 batches an ensemble that shares the tables (BASELINE config 2): every 2-D field
    gets `members` copies stacked along the layer axis; columns are members*n.
    """

    def __init__(self, A, O, S, kmax, ncmax=1, index_h2ovap=1, tabs=None, consts=None, members=1,
                 fast=False, device=None, layout=None, structured=True, ops=None, order_as=1, lon_mode=1):
        """layout (sharded runs, sharding.py): {"A"|"S"|"O": (n_own, n_ext, off)} -- cells this rank
        owns, cells of its source buffers (own + halo rows) and where the owned cells start in them;
        A/O/S are then objects with .im/.jm/.n of the LOCAL band.
        order_as: accuracy order of the conservative A->S table (gmapgen's interp_order_AS: 1, or 2 = the reference
        tool's default, three source rows per stencil); lon_mode: 0 = the reference generator's equal-longitude rule,
        1 = generalised longitude overlap (needed when ocean and atmosphere longitudes differ)."""
        import torch
        self.torch = torch
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.A, self.O, self.S = A, O, S
        self.K, self.nc, self.iq, self.M = kmax, ncmax, index_h2ovap, members
        self.fast = bool(fast)
        from . import synthetic as syn
        c = consts or {}
        self.Grav = c.get("Grav", syn.GRAV); self.CpDry = c.get("CpDry", syn.CPDRY)
        self.GasRDry = c.get("GasRDry", syn.GASRDRY); self.DelTime = c.get("DelTime", syn.DELTIME)
        self.sig1 = c.get("Sig1", syn.SIG1)
        layout = layout or {"A": (A.n, A.n, 0), "S": (S.n, S.n, 0), "O": (O.n, O.n, 0)}
        self.layout = layout
        self.sharded = any(ext != own for own, ext, off in layout.values())
        assert not (self.sharded and members > 1), "ensembles shard by member, not by latitude band"
        self.ops = {}
        self.nnz = {}
        if ops is not None:            # prebuilt operators (sharding.py: band operators straight from the grid axes)
            self.ops = dict(ops)
            self.nnz = {k: op.nnz for k, op in self.ops.items()}
        elif tabs is None and not self.sharded and structured:
            # straight from the grids (dccm_remap_create_jones99 / _bilinear): no table is built on the host;
            # the ocean-side pairs with different longitudes come back in separable form (kind 2)
            g = {"a": A, "s": S, "o": O}
            for key in ("as", "sa", "os", "so"):
                for kind in ("cons", "bil"):
                    order = order_as if (key == "as" and kind == "cons") else 1
                    self.ops[f"{key}_{kind}"] = RemapOperator.from_grids(g[key[0]], g[key[1]], kind == "cons", order, lon_mode)
            self.nnz = {k: op.nnz for k, op in self.ops.items()}
        else:
            tabs = tabs or build_tables(A, O, S, order_as=order_as, lon_mode=lon_mode)
            for key, (send_i, recv_i, coef) in tabs.items():
                s, d = key[0].upper(), key[1].upper()
                gx = {"A": A.im, "S": S.im, "O": O.im}
                self.ops[key] = RemapOperator(send_i, recv_i, coef, layout[s][1], layout[d][0],
                                              gnxs=gx[s] if structured else 0, gnxr=gx[d] if structured else 0)
                self.nnz[key] = len(coef)
        nA, nS, nO, M = A.n, S.n, O.n, members
        nAx, nSx, nOx = layout["A"][1], layout["S"][1], layout["O"][1]
        self.offA, self.offS, self.offO = layout["A"][2], layout["S"][2], layout["O"][2]
        # ensemble members are extra columns for the column solve: imax*jmax = M*nA
        self.vdiff = SfcImplicitCoupling(M * A.im, A.jm, kmax, ncmax, index_h2ovap,
                                         self.Grav, self.CpDry, self.GasRDry, self.DelTime, fast=fast)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.dev)
        # ATM side
        self.col_in = {k: None for k in IN_ORDER}
        self.tend = {"DUDt": z(kmax, M * nA), "DVDt": z(kmax, M * nA), "DTempDt": z(kmax, M * nA),
                     "DQMixDt": z(ncmax, kmax, M * nA)}
        # layer-major send/recv buffers; member m of layer l is row l*M + m
        self.a2s_bil = z(13 * M, nAx); self.a2s_cons = z(4 * M, nAx)
        self.o2s_bil = z(2 * M, nOx); self.o2s_cons = z(3 * M, nOx)
        self.s_bil = z(13 * M, nS); self.s_cons = z(4 * M, nS)
        self.s_obil = z(3 * M, nS)        # SfcTemp slots 1,2 (in) + slot 3 (out)
        self.s_ocons = z(4 * M, nS)       # SIceCon, SfcAlbedo slots 1,2 (in) + slot 3 (out)
        self.sfc_out = {k: z(3 * M, nS) for k in dsfcm.OUT3}
        self.sfc_out["DelVarImplCPL"] = z(4 * M, nS)
        self.s2a = z(9 * M, nSx); self.s2o = z(12 * M, nSx)
        self.a_recv = z(9 * M, nA); self.o_recv = z(12 * M, nO)
        if self.sharded:
            self.vdiff.set_coef_stride(nAx)
        self.launches = 0

    # -- inputs -----------------------------------------------------------------------------
    def set_inputs(self, col_in, atm_sfc, ocn_sfc):
        """col_in: dict IN_ORDER -> tensors with M*nA columns; atm_sfc/ocn_sfc: dicts of (M, n) tensors."""
        M = self.M
        self.col_in = col_in
        a0, a1 = self.offA, self.offA + self.A.n
        o0, o1 = self.offO, self.offO + self.O.n
        for l, name in enumerate(("WindU", "WindV", "SfcAirTemp", "QVap1", "SfcPress")):
            self.a2s_bil[l * M:(l + 1) * M, a0:a1] = atm_sfc[name]
        for l, name in enumerate(A2S_CONS):
            self.a2s_cons[l * M:(l + 1) * M, a0:a1] = atm_sfc[name]
        for l, name in enumerate(O2S_BIL):
            self.o2s_bil[l * M:(l + 1) * M, o0:o1] = ocn_sfc[name]
        for l, name in enumerate(O2S_CONS):
            self.o2s_cons[l * M:(l + 1) * M, o0:o1] = ocn_sfc[name]

    # -- the step ---------------------------------------------------------------------------
    def forward(self):
        M, nA = self.M, self.A.n
        out = dict(self.tend)
        # Coef1/Coef2 go straight into the A->S send buffer (layers 5..8 and 9..12); for M > 1 the
        # solver's (4, M*nA) slot-major layout is exactly rows [5M, 9M) of the (13M, nA) buffer.
        if self.sharded:      # strided slots inside the wider send buffer (dccm_vdiff_set_coef_stride)
            out["ImplCplCoef1"] = self.a2s_bil[5:9].view(-1)[self.offA:]
            out["ImplCplCoef2"] = self.a2s_bil[9:13].view(-1)[self.offA:]
        else:
            out["ImplCplCoef1"] = self.a2s_bil[5 * M:9 * M].view(4, M * nA)
            out["ImplCplCoef2"] = self.a2s_bil[9 * M:13 * M].view(4, M * nA)
        self.vdiff.forward_device(self.col_in, out)
        self.launches += 1 if self.fast else 2          # reference order: the solve + its (normally empty) IEEE redo kernel

    def remap_to_sfc(self):
        assert not self.sharded, "the unfused surface step is single-GPU only"
        M = self.M
        self.ops["as_bil"].apply(self.a2s_bil, self.s_bil)
        self.ops["as_cons"].apply(self.a2s_cons, self.s_cons)
        self.ops["os_bil"].apply(self.o2s_bil, self.s_obil[:2 * M])
        self.ops["os_cons"].apply(self.o2s_cons, self.s_ocons[:3 * M])
        self.launches += 4

    def bulk(self):
        M, nS = self.M, self.S.n
        # with M members the 3-slot arrays are (3M, nS): slot stride = M*nS, columns = M*nS
        v = lambda t, lo, hi: t[lo * M:hi * M].view(-1)
        f = dict(self.sfc_out)
        f = {k: t.view(-1) for k, t in f.items()}
        f.update(WindU=v(self.s_bil, 0, 1), WindV=v(self.s_bil, 1, 2), SfcAirTemp=v(self.s_bil, 2, 3),
                 QVap1=v(self.s_bil, 3, 4), SfcPress=v(self.s_bil, 4, 5),
                 ImplCplCoef1=v(self.s_bil, 5, 9), ImplCplCoef2=v(self.s_bil, 9, 13),
                 LDwRFlx=v(self.s_cons, 0, 1), SDwRFlx=v(self.s_cons, 1, 2),
                 SfcTemp=v(self.s_obil, 0, 3), SIceCon=v(self.s_ocons, 0, 1), SfcAlbedo=v(self.s_ocons, 1, 4),
                 SfcHeight=None)
        dsfcm.bulkflux_device(M * nS, 1, f, self.sig1, slot_stride=M * nS)
        self.launches += 1

    def pack_sfc(self):
        """put-side selection of the reference glue (ref sfc/dccm_sfc_mod.f90:764-784)."""
        M = self.M
        o = self.sfc_out
        sl = lambda t, n: t[(n - 1) * M:n * M]
        a = self.s2a
        a[0 * M:1 * M] = sl(o["LUwRFlx"], 3); a[1 * M:2 * M] = sl(o["SUwRFlx"], 3)
        a[2 * M:3 * M] = sl(o["SenHFlx"], 3); a[3 * M:4 * M] = sl(o["QVapMFlx"], 3)
        a[4 * M:5 * M] = self.s_ocons[3 * M:4 * M]
        a[5 * M:9 * M] = o["DelVarImplCPL"]
        b = self.s2o
        b[0 * M:1 * M] = sl(o["SfcHFlx_ns"], 1); b[1 * M:2 * M] = sl(o["SfcHFlx_sr"], 1)
        b[2 * M:3 * M] = self.s_cons[3 * M:4 * M]; b[3 * M:4 * M] = self.s_cons[2 * M:3 * M]
        b[4 * M:5 * M] = sl(o["QVapMFlx"], 1)
        self.torch.neg(sl(o["WindStressX"], 3), out=b[5 * M:6 * M])
        self.torch.neg(sl(o["WindStressY"], 3), out=b[6 * M:7 * M])
        b[7 * M:8 * M] = sl(o["SfcHFlx_ns"], 2); b[8 * M:9 * M] = sl(o["SfcHFlx_sr"], 2)
        b[9 * M:10 * M] = sl(o["QVapMFlx"], 2)
        b[10 * M:11 * M] = sl(o["DSfcHFlxDTs"], 1); b[11 * M:12 * M] = sl(o["DSfcHFlxDTs"], 2)

    def sfc_fused(self, store_full=False):
        """remap_to_sfc + bulk + pack_sfc in ONE kernel (dccm_sfc_exchange_device): the 22 remapped
        input layers stay in registers; only the 21 put-side layers are stored."""
        import ctypes as C
        full = None
        if store_full:
            f = L.SfcFields()
            for n in L.SfcFields._names:
                setattr(f, n, None)
            for k, t in self.sfc_out.items():
                setattr(f, k, t.data_ptr())
            f.SfcTemp = self.s_obil.data_ptr()
            f.SfcAlbedo = self.s_ocons[self.M:].data_ptr()
            full = C.byref(f)
        o = self.ops
        L.check(L.lib().dccm_sfc_exchange_device(
            o["as_bil"]._h, o["as_cons"]._h, o["os_bil"]._h, o["os_cons"]._h,
            L.tptr(self.a2s_bil), L.tptr(self.a2s_cons), L.tptr(self.o2s_bil), L.tptr(self.o2s_cons),
            self.M, float(self.sig1), C.c_void_p(self.s2a.data_ptr() + 8 * self.offS),
            C.c_void_p(self.s2o.data_ptr() + 8 * self.offS), self.s2a.shape[1], full, L.current_stream()))
        self.launches += 2            # the surface kernel + the (normally empty) IEEE redo kernel

    def configure_sfc(self, staged=-1, min_blocks=-1):
        """form of the fused surface kernel for THIS exchange (dccm_sfc_exchange_config; per handle, not process-wide)"""
        L.check(L.lib().dccm_sfc_exchange_config(self.ops["as_bil"]._h, staged, min_blocks))

    def sfc_last_form(self):
        return int(L.lib().dccm_sfc_exchange_last_form(self.ops["as_bil"]._h))

    # Option: the S->O / S->I remaps (latency bound, half the SM's warp slots and registers unused) on a second stream next
    # to the S->A remaps and the backward solve, which does not need them; backward() joins the two streams again.
    # Measured neutral at config 5 (9.38 vs 9.37-9.40 ms per exchange: the block scheduler hardly mixes the grids), so off.
    overlap_remaps = False

    def _fork_side(self):
        torch = self.torch
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_rside", None) is None:
            self._rside = torch.cuda.Stream(self.dev)
        ev = torch.cuda.Event(); ev.record(main)
        self._rside.wait_event(ev)
        return self._rside

    def _join_side(self):
        ev = getattr(self, "_rpending", None)
        if ev is not None:
            self.torch.cuda.current_stream(self.dev).wait_event(ev)
            self._rpending = None

    def _remaps_to_ocean(self):
        M = self.M
        self.ops["so_cons"].apply(self.s2o[:10 * M], self.o_recv[:10 * M])
        self.ops["so_bil"].apply(self.s2o[10 * M:], self.o_recv[10 * M:])

    def _remaps_to_atmosphere(self):
        M = self.M
        self.ops["sa_cons"].apply(self.s2a[:4 * M], self.a_recv[:4 * M])
        self.ops["sa_bil"].apply(self.s2a[4 * M:], self.a_recv[4 * M:])

    def remap_from_sfc(self):
        if self.overlap_remaps:
            side = self._fork_side()
            with self.torch.cuda.stream(side):
                self._remaps_to_ocean()
                self._rpending = self.torch.cuda.Event(); self._rpending.record(side)
            self._remaps_to_atmosphere()
        else:
            self._remaps_to_atmosphere()
            self._remaps_to_ocean()
        self.launches += 4

    def backward(self):
        M, nA = self.M, self.A.n
        lvl1 = self.a_recv[5 * M:9 * M].view(4, M * nA)
        self.vdiff.backward_device(self.tend, lvl1)
        self.launches += 1
        self._join_side()

    # -- latitude-slab pipeline: the issue-bound surface kernel next to the HBM-bound forward solve ------------------
    def _slab_plan(self, nslab):
        """S-row slabs cut at ATM cell edges, and for each the ATM rows whose forward solve it reads (sharding.BandPlan
        used for its row bookkeeping only -- everything stays in this object's buffers, no halo moves)"""
        key = ("slabs", nslab)
        if getattr(self, "_slab_cache", None) is None or self._slab_cache[0] != key:
            from .sharding import BandPlan
            plan = BandPlan(self.A, self.O, self.S, nslab)
            self._slab_cache = (key, [(plan.bands["A"][k], plan.bands["S"][k], plan.ext["A"][k]) for k in range(nslab)])
        return self._slab_cache[1]

    def _sfc_rows(self, row0, row1):
        import ctypes as C
        o = self.ops
        if getattr(self, "_whole_seg", None) is None:
            def seg(t):
                s = L.SrcSeg()
                s.lo = s.own = s.hi = t.data_ptr()
                s.b0, s.b1 = 0, 2 ** 63 - 1
                return s
            self._whole_seg = [seg(t) for t in (self.a2s_bil, self.a2s_cons, self.o2s_bil, self.o2s_cons)]
        sg = self._whole_seg
        L.check(L.lib().dccm_sfc_exchange_rows_device(
            o["as_bil"]._h, o["as_cons"]._h, o["os_bil"]._h, o["os_cons"]._h,
            C.byref(sg[0]), C.byref(sg[1]), C.byref(sg[2]), C.byref(sg[3]), 0, 0,
            self.M, float(self.sig1), C.c_void_p(self.s2a.data_ptr()), C.c_void_p(self.s2o.data_ptr()),
            self.s2a.shape[1], None, int(row0), int(row1), L.current_stream()))
        self.launches += 2

    def step_pipelined(self, nslab=16):
        """One exchange with the forward solve cut into `nslab` latitude slabs on the current stream and the fused
        surface kernel of every slab launched on a second, high-priority stream as soon as the ATM rows it reads are
        solved: the surface kernel is instruction-issue bound, the column solve HBM bound, so run side by side they
        share the SMs instead of taking turns.  Same kernels on row ranges -> same bits as step()."""
        torch = self.torch
        assert not self.sharded and self.M == 1, "slab pipelining: single-GPU, single-member exchanges"
        slabs = self._slab_plan(nslab)
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.dev, priority=-1)
        side, im = self._side, self.A.im
        M, nA = self.M, self.A.n
        out = dict(self.tend)
        out["ImplCplCoef1"] = self.a2s_bil[5 * M:9 * M].view(4, M * nA)
        out["ImplCplCoef2"] = self.a2s_bil[9 * M:13 * M].view(4, M * nA)
        start = torch.cuda.Event(); start.record(main)
        side.wait_event(start)                      # the side stream joins the work (and any stream capture) here
        k_sfc = 0
        for k, ((a0, a1), _, _) in enumerate(slabs):
            self.vdiff.forward_cols_device(self.col_in, out, a0 * im, a1 * im)
            self.launches += 1 if self.fast else 2
            ev = None
            while k_sfc < nslab and (slabs[k_sfc][2][1] <= a1 or k == nslab - 1):
                if ev is None:
                    ev = torch.cuda.Event(); ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self._sfc_rows(*slabs[k_sfc][1])
                k_sfc += 1
        done = torch.cuda.Event(); done.record(side)
        main.wait_event(done)
        self.remap_from_sfc()
        self.backward()

    def halo_to_sfc(self):
        """hook: sharded runs exchange the halo rows of the ATM/OCN send buffers here"""

    def halo_from_sfc(self):
        """hook: sharded runs exchange the halo rows of the SFC send buffers here"""

    def step(self, fused=True):
        self.forward()
        self.halo_to_sfc()
        if fused:
            self.sfc_fused()
        else:
            self.remap_to_sfc()
            self.bulk()
            self.pack_sfc()
        self.halo_from_sfc()
        self.remap_from_sfc()
        self.backward()

    # -- accounting (SURVEY.md 8d / BASELINE.md section 4) --------------------------------------
    def algorithmic_bytes(self, n_out=42):
        nA, nS, nO, K, nc, M = self.A.n, self.S.n, self.O.n, self.K, self.nc, self.M

        def remap(key, D, nsrc, ndst):
            return 12 * self.nnz[key] + 4 * (ndst + 1) + 8 * D * M * (nsrc + ndst)
        # NOTE: the accounting keeps the CSR table bytes even for tables stored as zonal stencils
        # (kind 1), so "algorithmic bytes" stay comparable across builds (SURVEY 8d formula).

        b = {}
        b["fwd"] = 8 * M * nA * ((9 + nc) * (K + 1) + 2 * K + 3 * K + (3 + nc) * K + 8)
        b["bwd"] = 8 * M * nA * (3 * K + (3 + nc) * K + 4 + (3 + nc) * K)
        b["remap_to_sfc"] = (remap("as_bil", 13, nA, nS) + remap("as_cons", 4, nA, nS)
                             + remap("os_bil", 2, nO, nS) + remap("os_cons", 3, nO, nS))
        b["bulk"] = 8 * (20 + n_out) * M * nS
        b["remap_from_sfc"] = (remap("sa_cons", 4, nS, nA) + remap("sa_bil", 5, nS, nA)
                               + remap("so_cons", 10, nS, nO) + remap("so_bil", 2, nS, nO))
        # fused surface step: tables + source layers in, the 21 put-side layers out
        b["sfc_fused"] = (12 * (self.nnz["as_bil"] + self.nnz["as_cons"] + self.nnz["os_bil"] + self.nnz["os_cons"])
                          + 4 * 4 * (nS + 1) + 8 * M * (17 * nA + 5 * nO) + 8 * 21 * M * nS)
        b["total"] = b["fwd"] + b["bwd"] + b["remap_to_sfc"] + b["bulk"] + b["remap_from_sfc"]
        b["total_fused"] = b["fwd"] + b["bwd"] + b["sfc_fused"] + b["remap_from_sfc"]
        return b

    def moved_bytes(self):
        """bytes the kernels of one (fused) exchange actually move: like algorithmic_bytes()["total_fused"], but a table
        stored as zonal stencils (kind 1) or in separable form (kind 2) contributes no 12*nnz / row-pointer bytes --
        its few KB of stencil / factor lists stay in L1/L2."""
        b = self.algorithmic_bytes()
        nA, nS, nO = self.A.n, self.S.n, self.O.n
        ndst = {"as": nS, "os": nS, "sa": nA, "so": nO}
        tab = sum(12 * self.nnz[k] + 4 * (ndst[k[:2]] + 1) for k, op in self.ops.items() if op.kind != 0)
        return b["total_fused"] - tab

    def remapped_cell_fields(self):
        nA, nS, nO, M = self.A.n, self.S.n, self.O.n, self.M
        return M * (22 * nS + 9 * nA + 12 * nO)


# ---------------------------------------------------------------------------------------------
# Legacy 2-component topology: what the shipped exp/ configurations run.
# The atmosphere computes the surface fluxes itself and ships 12 flux fields straight to the ocean;
# 4 fields come back (ref common/mod_common_params.f90:76-175, ocn/mod_ocn.f90:615-669,
# atm/mod_atm.f90:427-442).  Same remap operator (K1), other field lists, no bulk flux / solve.
A2O_CONS = ["WindStressX", "WindStressY", "LDwRFlx", "SDwRFlx", "LUwRFlx", "SUwRFlx", "LatHFlx", "SenHFlx",
            "RainFall", "SnowFall"]                      # GMAPTAG_ATM2D_OCN2D_CONSERVE
A2O_BIL = ["DSfcHFlxDTs", "SfcAirTemp"]                  # GMAPTAG_ATM2D_OCN2D
O2A_CONS = ["SfcTemp", "SfcAlbedo", "SfcEngyFlxMod"]
O2A_BIL = ["SfcSnow"]


class LegacyExchange:
    """A<->O exchange of the 2-component mode on the device: 12 + 4 layers through four tables."""

    def __init__(self, A, O, order_ao=1, lon_mode=1, device=None, tabs=None):
        import torch
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.A, self.O = A, O
        if tabs is None:
            tabs = {"ao_cons": tables.gen_table_jones99(A, O, order_ao, lon_mode).index(A.im, O.im),
                    "ao_bil": tables.gen_table_bilinear(A, O, lon_mode).index(A.im, O.im),
                    "oa_cons": tables.gen_table_jones99(O, A, 1, lon_mode).index(O.im, A.im),
                    "oa_bil": tables.gen_table_bilinear(O, A, lon_mode).index(O.im, A.im)}
        self.tabs = tabs
        g = {"a": A, "o": O}
        self.ops = {k: RemapOperator(*t, g[k[0]].n, g[k[1]].n, gnxs=g[k[0]].im, gnxr=g[k[1]].im) for k, t in tabs.items()}
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.dev)
        self.a2o = z(12, A.n); self.o_recv = z(12, O.n)
        self.o2a = z(4, O.n); self.a_recv = z(4, A.n)

    def step(self):
        self.ops["ao_cons"].apply(self.a2o[:10], self.o_recv[:10])
        self.ops["ao_bil"].apply(self.a2o[10:], self.o_recv[10:])
        self.ops["oa_cons"].apply(self.o2a[:3], self.a_recv[:3])
        self.ops["oa_bil"].apply(self.o2a[3:], self.a_recv[3:])
