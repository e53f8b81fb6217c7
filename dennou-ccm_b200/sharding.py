"""Row-block sharding of the exchange step over N GPUs.

Every grid is cut into N latitude bands; rank g owns band g of the ATM, SFC and OCN grids
(contiguous blocks of the i-fastest linear index, ref common/field_def.f90:622-629), the
destination rows of every mapping table that fall into its bands (the reference partitions the
remap by the receiving component's local cells, ref common/interpolation_data_latlon_mod.f90:
140-151), and the columns of the column solve / bulk flux in them.  Cuts are made at ATM cell
edges; those are also SFC cell edges (the exchange grid merges ATM and OCN edges,
ref tool/gmapgen/gmapgen_main.f90:376-377), so SFC bands cover ATM bands exactly and the
conservative ATM<->SFC remaps need no halo; bilinear stencils and the OCN rows that straddle a
cut need the neighbouring band's boundary rows.

Per coupling interval the only inter-GPU traffic is that halo: a few source rows of each send
buffer, exchanged with the two neighbouring ranks (NCCL send/recv grouped in one batch) -- for
the T1279 <-> 0.1 deg case about 1 MB per rank instead of the 1 GB an allgather of the source
fields would move.

The plan (bands, halo ranges, local tables) is pure host logic and identical on every rank, so
no metadata is ever communicated.
"""
import numpy as np

from . import tables
from .exchange import SurfaceExchange


class Local:
    """A latitude band of a grid, exposing what SurfaceExchange needs (.im/.jm/.n)."""

    def __init__(self, grid, j0, j1):
        self.grid, self.j0, self.j1 = grid, j0, j1
        self.im, self.jm = grid.im, j1 - j0
        self.n = self.im * self.jm


def _edges(g):
    """latitudes of the cell edges as the generators reconstruct them
    (ref common/grid_mapping_util_jones99.f90:147-151)."""
    v = np.empty(g.jm + 1)
    v[0] = -np.pi / 2
    for j in range(1, g.jm):
        v[j] = np.arcsin(g.y_LatWt[j - 1] + np.sin(v[j - 1]))
    v[g.jm] = np.pi / 2
    return v


class BandPlan:
    """bands[grid][rank] = (j0, j1) owned rows; ext[grid][rank] = (e0, e1) rows of the source
    buffers (own + halo); tables are generated per rank for its destination rows only."""

    GRIDS = ("A", "S", "O")
    TABLES = ("as", "os", "sa", "so")        # source grid, destination grid

    def __init__(self, A, O, S, world, order_as=1, lon_mode=1):
        self.A, self.O, self.S, self.world = A, O, S, world
        self.grid = {"A": A, "S": S, "O": O}
        self.order_as, self.lon_mode = order_as, lon_mode
        if world > min(A.jm, O.jm):
            raise ValueError(f"{world} ranks for {A.jm} ATM / {O.jm} OCN rows: bands would be empty")
        eA, eS = _edges(A), _edges(S)
        cutA = [int(round(g * A.jm / world)) for g in range(world + 1)]
        # the SFC row whose south edge is the ATM cut edge (exact when JMA != JMO; when the
        # exchange grid IS the ATM grid the indices coincide)
        cutS = [0] + [int(np.argmin(np.abs(eS - eA[c]))) for c in cutA[1:-1]] + [S.jm]
        # OCN rows go to the band holding their centre
        cutO = [0] + [int(np.searchsorted(O.y_Lat, eA[c])) for c in cutA[1:-1]] + [O.jm]
        for cuts, name in ((cutA, "ATM"), (cutS, "SFC"), (cutO, "OCN")):
            if any(b <= a for a, b in zip(cuts[:-1], cuts[1:])):
                raise ValueError(f"empty {name} band with {world} ranks")
        self.bands = {"A": list(zip(cutA[:-1], cutA[1:])), "S": list(zip(cutS[:-1], cutS[1:])),
                      "O": list(zip(cutO[:-1], cutO[1:]))}
        # source-row extent needed by each rank = union over the tables reading that grid
        self.ext = {g: [list(self.bands[g][r]) for r in range(world)] for g in self.GRIDS}
        for r in range(world):
            for key in self.TABLES:
                s, d = key[0].upper(), key[1].upper()
                for kind in ("cons", "bil"):
                    lo, hi = self._src_rows(key, kind, r)
                    self.ext[s][r][0] = min(self.ext[s][r][0], lo)
                    self.ext[s][r][1] = max(self.ext[s][r][1], hi)
        for g in self.GRIDS:
            for r in range(world):
                e0, e1 = self.ext[g][r]
                lo_ok = e0 >= (self.bands[g][r - 1][0] if r > 0 else 0)
                hi_ok = e1 <= (self.bands[g][r + 1][1] if r < world - 1 else self.grid[g].jm)
                if not (lo_ok and hi_ok):
                    raise ValueError(f"rank {r}: halo of grid {g} reaches beyond the neighbouring band "
                                     f"(bands too thin for {world} ranks)")

    # -- tables -------------------------------------------------------------------------
    def _gen(self, key, kind, rows):
        """table entries (GLOBAL 1-based indices) for destination rows [rows[0], rows[1])"""
        s, d = self.grid[key[0].upper()], self.grid[key[1].upper()]
        if kind == "cons":
            order = self.order_as if key == "as" else 1
            t = tables.gen_table_jones99(s, d, order, self.lon_mode, rows=rows)
        else:
            t = tables.gen_table_bilinear(s, d, self.lon_mode, rows=rows)
        return t.index(s.im, d.im)

    def _src_rows(self, key, kind, rank):
        """source rows referenced by the band: the stencils move monotonically with the destination
        row, so the first and last destination rows give the extent."""
        j0, j1 = self.bands[key[1].upper()][rank]
        im = self.grid[key[0].upper()].im
        lo, hi = None, None
        for rows in {(j0, j0 + 1), (j1 - 1, j1)}:
            send = self._gen(key, kind, rows)[0]
            if len(send):
                a, b = int((send.min() - 1) // im), int((send.max() - 1) // im) + 1
                lo, hi = (a if lo is None else min(lo, a)), (b if hi is None else max(hi, b))
        return self.bands[key[0].upper()][rank] if lo is None else (lo, hi)

    def local_tables(self, rank):
        """dict like exchange.build_tables(), indices LOCAL to the rank: destination rows relative
        to the owned band, source cells relative to the rank's source buffer (ext rows)."""
        out = {}
        for key in self.TABLES:
            s, d = key[0].upper(), key[1].upper()
            for kind in ("cons", "bil"):
                send, recv, coef = self._gen(key, kind, self.bands[d][rank])
                soff = self.ext[s][rank][0] * self.grid[s].im
                doff = self.bands[d][rank][0] * self.grid[d].im
                out[f"{key}_{kind}"] = ((send - soff).astype(np.int32), (recv - doff).astype(np.int32), coef)
        return out

    def local_operators(self, rank):
        """the eight operators of the rank's bands straight from the grid axes (RemapOperator.from_grids with rows /
        src_rows): zonal stencils where the longitudes agree, separable factors where they differ -- no table is
        generated on the host and none is read by the kernels; same bits as the rows of the unsharded operators."""
        from .interpolation_data_latlon_mod import RemapOperator
        out = {}
        for key in self.TABLES:
            s, d = key[0].upper(), key[1].upper()
            for kind in ("cons", "bil"):
                order = self.order_as if (key == "as" and kind == "cons") else 1
                out[f"{key}_{kind}"] = RemapOperator.from_grids(self.grid[s], self.grid[d], kind == "cons", order, self.lon_mode,
                                                                rows=tuple(self.bands[d][rank]), src_rows=tuple(self.ext[s][rank]))
        return out

    def layout(self, rank):
        lay = {}
        for g in self.GRIDS:
            im = self.grid[g].im
            (j0, j1), (e0, e1) = self.bands[g][rank], self.ext[g][rank]
            lay[g] = ((j1 - j0) * im, (e1 - e0) * im, (j0 - e0) * im)
        return lay

    def local_grids(self, rank):
        return tuple(Local(self.grid[g], *self.bands[g][rank]) for g in ("A", "O", "S"))

    # -- halo messages ------------------------------------------------------------------
    def halo_messages(self, g, rank):
        """[(peer, 'send'|'recv', c0, c1)] with c0:c1 cell ranges inside rank's OWN source buffer of
        grid g.  Lower-neighbour messages come first on both sides, so sends and receives pair up."""
        im, w = self.grid[g].im, self.world
        (j0, j1), (e0, e1) = self.bands[g][rank], self.ext[g][rank]
        cell = lambda j: (j - e0) * im
        msgs = []
        if rank > 0:
            pe0, pe1 = self.ext[g][rank - 1]
            if pe1 > j0:                                    # rows of mine the lower neighbour reads
                msgs.append((rank - 1, "send", cell(j0), cell(pe1)))
            if e0 < j0:
                msgs.append((rank - 1, "recv", cell(e0), cell(j0)))
        if rank < w - 1:
            pe0, pe1 = self.ext[g][rank + 1]
            if pe0 < j1:
                msgs.append((rank + 1, "send", cell(pe0), cell(j1)))
            if e1 > j1:
                msgs.append((rank + 1, "recv", cell(j1), cell(e1)))
        return msgs

    def halo_bytes(self, rank, layers):
        """bytes a rank receives per exchange (layers: dict grid -> number of layers)."""
        return sum(8 * layers[g] * (c1 - c0) for g in self.GRIDS
                   for peer, kind, c0, c1 in self.halo_messages(g, rank) if kind == "recv")


class HaloExchanger:
    """Halo rows of several layer-major send buffers, exchanged with the two neighbouring ranks.

    All layers of all buffers that go to one neighbour travel in ONE message (one staging tensor
    per peer and direction, allocated once), so a phase is at most 2 sends + 2 receives grouped in
    a single NCCL launch (gloo on CPU tensors in the tests), however many fields are coupled."""

    def __init__(self, bufs_msgs, dist):
        self.dist = dist
        self.items = []                      # (peer, kind, staging, [(buf, c0, c1, offset)])
        by = {}
        for buf, msgs in bufs_msgs:
            for peer, kind, c0, c1 in msgs:
                by.setdefault((peer, kind), []).append((buf, c0, c1))
        # lower neighbour first, sends and receives in the same order on both sides
        for (peer, kind) in sorted(by, key=lambda k: (k[0], k[1] != "send")):
            parts, n = [], 0
            for buf, c0, c1 in by[(peer, kind)]:
                parts.append((buf, c0, c1, n))
                n += buf.shape[0] * (c1 - c0)
            stage = parts[0][0].new_empty(n)
            self.items.append((peer, kind, stage, parts))
        self.ops = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, stage, peer)
                    for peer, kind, stage, parts in self.items] if dist is not None else []
        self.nbytes_recv = sum(8 * st.numel() for peer, kind, st, parts in self.items if kind == "recv")

    def __call__(self):
        if not self.items:
            return
        for peer, kind, stage, parts in self.items:
            if kind == "send":
                for buf, c0, c1, o in parts:
                    stage[o:o + buf.shape[0] * (c1 - c0)].view(buf.shape[0], c1 - c0).copy_(buf[:, c0:c1])
        for req in self.dist.batch_isend_irecv(self.ops):
            req.wait()
        for peer, kind, stage, parts in self.items:
            if kind == "recv":
                for buf, c0, c1, o in parts:
                    buf[:, c0:c1] = stage[o:o + buf.shape[0] * (c1 - c0)].view(buf.shape[0], c1 - c0)


class AllGatherHalo(HaloExchanger):
    """The same halo rows moved by ONE all-gather per phase (the collective north_star names: "NCCL-over-NVLink
    allgather of the source field halo each coupling interval").  Every rank contributes one fixed-size block
    [rows for its lower neighbour | rows for its upper neighbour] -- the boundary rows of all layers of all buffers,
    a few MB at most, never the fields -- and picks its two neighbours' segments out of the gathered blocks.
    Slot sizes are the maxima over the ranks (one MAX all-reduce at construction; unused tails travel as padding)."""

    def __init__(self, bufs_msgs, dist, rank, world):
        super().__init__(bufs_msgs, None)
        self.dist, self.rank, self.world = dist, rank, world
        torch = __import__("torch")
        size = {"lo": 0, "hi": 0}
        for peer, kind, stage, parts in self.items:
            if kind == "send":
                size["lo" if peer < rank else "hi"] = stage.numel()
        like = bufs_msgs[0][0]               # every rank takes part in the collective, with or without rows of its own
        t = torch.tensor([size["lo"], size["hi"]], dtype=torch.int64, device=like.device)
        if dist is not None and world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.n_lo, self.n_hi = int(t[0]), int(t[1])
        n = self.n_lo + self.n_hi
        if n > 0:
            self.block = like.new_zeros(n)
            self.gathered = like.new_empty(world * n)
        else:                                # no rank has halo rows for these buffers: nothing to do anywhere
            self.block = self.gathered = None

    def __call__(self):
        if self.block is None:
            return
        n = self.n_lo + self.n_hi
        for peer, kind, stage, parts in self.items:
            if kind == "send":
                base = 0 if peer < self.rank else self.n_lo
                for buf, c0, c1, o in parts:
                    m = buf.shape[0] * (c1 - c0)
                    self.block[base + o:base + o + m].view(buf.shape[0], c1 - c0).copy_(buf[:, c0:c1])
        self.dist.all_gather_into_tensor(self.gathered, self.block)
        for peer, kind, stage, parts in self.items:
            if kind == "recv":
                # what the lower neighbour sent upwards sits in the "hi" slot of its block, and vice versa
                base = peer * n + (self.n_lo if peer < self.rank else 0)
                for buf, c0, c1, o in parts:
                    m = buf.shape[0] * (c1 - c0)
                    buf[:, c0:c1] = self.gathered[base + o:base + o + m].view(buf.shape[0], c1 - c0)


def exchange_halo(bufs_msgs, rank, dist, mode="sendrecv", world=None):
    """one-shot form of HaloExchanger / AllGatherHalo (tests)"""
    if mode == "allgather":
        AllGatherHalo(bufs_msgs, dist, rank, world)()
    else:
        HaloExchanger(bufs_msgs, dist)()


class ShardedExchange(SurfaceExchange):
    """SurfaceExchange on rank `rank` of `world`: local bands + halo exchange around the surface step."""

    def __init__(self, A, O, S, kmax, ncmax=1, index_h2ovap=1, rank=0, world=1, plan=None, dist=None,
                 halo="sendrecv", band_tables=False, order_as=1, lon_mode=1, **kw):
        """halo: "sendrecv" = grouped NCCL send/recv with the two neighbours, "allgather" = one all-gather of every
        rank's boundary rows per phase (AllGatherHalo).  band_tables: generate the band's tables on the host and
        keep them as CSR (the round-1 form; tests compare it with the table-free band operators)"""
        self.plan = plan or BandPlan(A, O, S, world, order_as=order_as, lon_mode=lon_mode)
        self.rank, self.world, self.dist = rank, world, dist
        lA, lO, lS = self.plan.local_grids(rank)
        if band_tables:
            super().__init__(lA, lO, lS, kmax, ncmax, index_h2ovap, tabs=self.plan.local_tables(rank),
                             layout=self.plan.layout(rank), **kw)
        else:
            super().__init__(lA, lO, lS, kmax, ncmax, index_h2ovap, ops=self.plan.local_operators(rank),
                             layout=self.plan.layout(rank), **kw)
        mA, mO, mS = (self.plan.halo_messages(g, rank) for g in ("A", "O", "S"))
        bufs_in = [(self.a2s_bil, mA), (self.a2s_cons, mA), (self.o2s_bil, mO), (self.o2s_cons, mO)]
        bufs_out = [(self.s2a, mS), (self.s2o, mS)]
        if halo == "allgather" and dist is not None:
            self.halo_in = AllGatherHalo(bufs_in, dist, rank, world)
            self.halo_out = AllGatherHalo(bufs_out, dist, rank, world)
        else:
            self.halo_in = HaloExchanger(bufs_in, dist)
            self.halo_out = HaloExchanger(bufs_out, dist)

    def halo_to_sfc(self):
        if self.world > 1:
            self.halo_in()

    def halo_from_sfc(self):
        if self.world > 1:
            self.halo_out()


class PeerShardedExchange(ShardedExchange):
    """Sharded exchange whose halo never moves: the surface kernel and the S->A / S->O remaps read
    the neighbouring bands' boundary rows IN PLACE from the neighbours' own send buffers, which are
    allocated as symmetric memory and peer-mapped over NVLink (torch.distributed._symmetric_memory).
    The remap and its halo "collective" are one kernel; per exchange the only inter-GPU operations
    besides those loads are two device-side barriers (after the forward solve, after the surface
    kernel), which also order the buffer reuse of the next exchange:

        forward -> B1 -> surface kernel (reads a2s/o2s rows of rank+-1) -> B2 -> remaps (read s2a/s2o rows
        of rank+-1) -> backward -> [next exchange] forward (overwrites a2s: every reader passed B2) ...

    Falls back to nothing: construct ShardedExchange (NCCL send/recv halo) where peer access is missing.
    """

    BUFS = (("a2s_bil", 13, "A"), ("a2s_cons", 4, "A"), ("o2s_bil", 2, "O"), ("o2s_cons", 3, "O"),
            ("s2a", 9, "S"), ("s2o", 12, "S"))

    def __init__(self, A, O, S, kmax, ncmax=1, index_h2ovap=1, rank=0, world=1, plan=None, dist=None,
                 sync="barrier", **kw):
        """sync: "barrier" = one all-rank device barrier per phase (default: one small kernel; 788-790 exchanges/s on
        8 GPUs), "neighbour" = pairwise signals with the two neighbouring ranks only (four small kernels; 782-783)"""
        import ctypes as C
        import torch
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        super().__init__(A, O, S, kmax, ncmax, index_h2ovap, rank=rank, world=world, plan=plan, dist=dist, **kw)
        assert self.M == 1
        self.sync = sync
        self._nbrs = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
        self.sharded = True                                 # rows are wider than the owned band
        plan = self.plan
        self.seg, self.rowlen, self._hdl = {}, {}, {}
        group = dist.group.WORLD
        # ONE symmetric allocation (one rendezvous: ~0.5 s each on 8 GPUs) holds the six send buffers, each starting on a
        # 128-byte boundary; row lengths are the maxima over the ranks so that every rank's layout is the same
        shapes, total = [], 0
        for name, nl, g in self.BUFS:
            im = plan.grid[g].im
            nmax = max((plan.ext[g][r][1] - plan.ext[g][r][0]) * im for r in range(world))
            nmax = (nmax + 1) & ~1                          # even: 16-byte aligned layer rows for the bulk copies
            shapes.append((name, nl, g, nmax, total))
            total += (nl * nmax + 15) & ~15
        flat = symm.empty((total,), dtype=torch.float64, device=self.dev)
        flat.zero_()
        hdl = symm.rendezvous(flat, group)
        self._flat = flat
        for name, nl, g, nmax, off in shapes:
            im = plan.grid[g].im
            t = flat[off:off + nl * nmax].view(nl, nmax)
            old = getattr(self, name)
            t[:, :old.shape[1]] = old                       # keep whatever set_inputs() already stored
            setattr(self, name, t)
            self._hdl[name] = hdl
            self.rowlen[g] = nmax
            (j0, j1), (e0, e1) = plan.bands[g][rank], plan.ext[g][rank]
            seg = L.SrcSeg()
            own = t.data_ptr()
            seg.own = own
            seg.b0, seg.b1 = (j0 - e0) * im, (j1 - e0) * im
            seg.lo = seg.hi = own
            if rank > 0 and e0 < j0:
                peer = hdl.get_buffer(rank - 1, (nl, nmax), torch.float64, off)
                seg.lo = peer.data_ptr() + 8 * (e0 - plan.ext[g][rank - 1][0]) * im
            if rank < world - 1 and e1 > j1:
                peer = hdl.get_buffer(rank + 1, (nl, nmax), torch.float64, off)
                seg.hi = peer.data_ptr() + 8 * (e0 - plan.ext[g][rank + 1][0]) * im
            self.seg[name] = seg
        self.vdiff.set_coef_stride(self.rowlen["A"])
        self._bar = self._hdl["a2s_bil"]
        torch.cuda.synchronize()
        dist.barrier()

    # Ordering between neighbours.  Phase 0 (after the forward solve): "my a2s / o2s rows are written" -- the surface
    # kernel may read the neighbours' boundary rows; phase 1 (after the surface kernel): "my s2a / s2o rows are written"
    # -- the remaps may read them.  The same two handshakes also order the buffer reuse of the next exchange: a rank
    # has seen its neighbour's phase-1 signal (sent after the neighbour's surface kernel, the last reader of this
    # rank's a2s rows) before its next forward solve overwrites a2s, and its neighbour's next phase-0 signal (sent after
    # the neighbour's remaps, the last readers of this rank's s2a rows) before its next surface kernel overwrites s2a.
    # Pairwise signals with the two neighbours are enough for that (sync="neighbour"; a 20 s timeout turns a lost signal
    # into a device trap); measured on 8 GPUs the single all-rank barrier kernel is 1 % faster than the four signal
    # kernels, so it stays the default.
    def _handshake(self, channel):
        if self.world <= 1:
            return
        if self.sync == "barrier":
            self._bar.barrier(channel=channel)
            return
        for r in self._nbrs:
            self._bar.put_signal(r, channel, 20000)
        for r in self._nbrs:
            self._bar.wait_signal(r, channel, 20000)

    def halo_to_sfc(self):
        self._handshake(0)

    def halo_from_sfc(self):
        self._handshake(1)

    def sfc_fused(self, store_full=False):
        import ctypes as C
        from . import _lib as L
        o, s = self.ops, self.seg
        L.check(L.lib().dccm_sfc_exchange_seg_device(
            o["as_bil"]._h, o["as_cons"]._h, o["os_bil"]._h, o["os_cons"]._h,
            C.byref(s["a2s_bil"]), C.byref(s["a2s_cons"]), C.byref(s["o2s_bil"]), C.byref(s["o2s_cons"]),
            self.rowlen["A"], self.rowlen["O"], 1, float(self.sig1),
            C.c_void_p(self.s2a.data_ptr() + 8 * self.offS), C.c_void_p(self.s2o.data_ptr() + 8 * self.offS),
            self.rowlen["S"], None, L.current_stream()))
        self.launches += 2            # the surface kernel + the (normally empty) IEEE redo kernel

    def remap_from_sfc(self):
        import ctypes as C
        from . import _lib as L

        def apply(key, name, row0, nrow, recv):
            seg = L.SrcSeg()
            base = self.seg[name]
            o = 8 * row0 * self.rowlen["S"]
            seg.lo, seg.own, seg.hi, seg.b0, seg.b1 = base.lo + o, base.own + o, base.hi + o, base.b0, base.b1
            L.check(L.lib().dccm_remap_apply_seg_device(self.ops[key]._h, C.byref(seg), self.rowlen["S"],
                                                        L.tptr(recv), recv.shape[1], recv.shape[0], nrow,
                                                        L.current_stream()))
        def to_ocean():
            apply("so_cons", "s2o", 0, 10, self.o_recv[:10])
            apply("so_bil", "s2o", 10, 2, self.o_recv[10:])
        if self.overlap_remaps:              # S->O / S->I next to S->A and the backward solve (exchange.py)
            side = self._fork_side()
            with self.torch.cuda.stream(side):
                to_ocean()
                self._rpending = self.torch.cuda.Event(); self._rpending.record(side)
        apply("sa_cons", "s2a", 0, 4, self.a_recv[:4])
        apply("sa_bil", "s2a", 4, 5, self.a_recv[4:])
        if not self.overlap_remaps:
            to_ocean()
        self.launches += 4
