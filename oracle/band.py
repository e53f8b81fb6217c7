"""CPU ORACLE (test infrastructure, NOT product code): the reference's whole exchange step on a LATITUDE BAND of the
grids, composed from oracle calls only, with the real data flow

    VDiffForward (ATM) -> interpolate_data A->S, O/I->S -> DSFCM_Util_SfcBulkFlux_Get -> put-side selection
    -> interpolate_data S->A, S->O/I -> level-1 update -> VDiffBackward (ATM)

(field lists and order: SURVEY.md Appendix A; ref sfc/dccm_sfc_mod.f90:449-466, :764-784, atm/dccm_atm_mod.f90:697-712,
:817-835, ocn/dccm_ocn_mod.f90:625-645).  Grids and mapping tables come from the oracle's own generators -- nothing of
the product package is imported, so bench.py's reference arm and cpu_baseline do not load libdccm_b200.so.

A band is given by the atmosphere rows [a0, a1) whose tendencies are wanted.  Working backwards through the tables
(only the lines of the destination rows in question are generated, orc_gen_*_rows) gives the exchange-grid rows the
S->A / S->O remaps read, then the atmosphere / ocean rows the A->S / O->S remaps read: the band is extended by those
few halo rows and every stage runs on exactly the rows it needs.  With a0 = 0, a1 = jm it is the whole grid.

Users: bench.py (cpu_baseline, --impl reference, the parity block: the GPU's outputs of the band rows against this, fed
with the SAME input bits) and tests/.
"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def lat_edges(wt):
    """cell edges from the quadrature weights, ref common/grid_mapping_util_jones99.f90:147-151"""
    v = np.empty(len(wt) + 1)
    v[0] = -np.pi / 2.0
    for j in range(1, len(wt)):
        v[j] = np.arcsin(wt[j - 1] + np.sin(v[j - 1]))
    v[-1] = np.pi / 2.0
    return v


def make_grids(orc, ima, jma, imo, jmo, ocean_regular):
    A = orc.gauss_grid(ima, jma)
    O = orc.regular_grid(imo, jmo) if ocean_regular else orc.gauss_grid(imo, jmo)
    return A, O, orc.exchange_grid(A, O)


class _Tab:
    """lines of one mapping table for destination rows [d0, d1), indices local to (source rows [s0, s1), those rows)"""

    def __init__(self, orc, kind, src, dst, rows, order=1, lon_mode=1):
        t = (orc.gen_bilinear(src, dst, lon_mode, rows=rows) if kind == "bil"
             else orc.gen_jones99(src, dst, order, lon_mode, rows=rows))
        self.send, self.recv, self.coef = t.to_index(src.im, dst.im)
        self.src_im, self.dst_im, self.rows = src.im, dst.im, rows
        self.smin = (int(self.send.min()) - 1) // src.im if t.n else 0
        self.smax = (int(self.send.max()) - 1) // src.im + 1 if t.n else 0

    def localize(self, s0):
        self.send = (self.send - s0 * self.src_im).astype(np.int32)
        self.recv = (self.recv - self.rows[0] * self.dst_im).astype(np.int32)
        self.n_dst = (self.rows[1] - self.rows[0]) * self.dst_im
        assert self.send.min() >= 1 and self.recv.min() >= 1 and self.recv.max() <= self.n_dst
        return self


class BandExchange:
    def __init__(self, orc, A, O, S, K, nc, a_rows, consts, iq=1, order_as=1, lon_mode=1, ranks=1):
        """consts: dict Grav, CpDry, GasRDry, DelTime, Sig1 (the DCPAM-side constants of the coupling solve and
        a_Sig1Info(1)).  ranks: the atmosphere's column solves run as that many independent column blocks, one host
        thread each -- the reference decomposes the atmosphere over MPI ranks (ref atm/dccm_atm_mod.f90:172-176)."""
        self.orc, self.A, self.O, self.S, self.K, self.nc, self.iq = orc, A, O, S, K, nc, iq
        self.c = consts
        a0, a1 = a_rows
        vA = lat_edges(A.y_LatWt)
        o0 = 0 if a0 == 0 else int(np.searchsorted(O.y_Lat, vA[a0]))
        o1 = O.jm if a1 == A.jm else int(np.searchsorted(O.y_Lat, vA[a1]))
        self.a_rows, self.o_rows = (a0, a1), (o0, o1)
        T = lambda kind, s, d, rows, order=1: _Tab(orc, kind, s, d, rows, order, lon_mode)
        sa = [T("cons", S, A, (a0, a1)), T("bil", S, A, (a0, a1))]
        so = [T("cons", S, O, (o0, o1)), T("bil", S, O, (o0, o1))] if o1 > o0 else []
        s0 = min(t.smin for t in sa + so)
        s1 = max(t.smax for t in sa + so)
        self.s_rows = (s0, s1)
        a_s = [T("bil", A, S, (s0, s1)), T("cons", A, S, (s0, s1), order_as)]
        o_s = [T("bil", O, S, (s0, s1)), T("cons", O, S, (s0, s1))]
        self.ae = (min(a0, min(t.smin for t in a_s)), max(a1, max(t.smax for t in a_s)))
        self.oe = (min(o0, min(t.smin for t in o_s)), max(o1, max(t.smax for t in o_s)))
        for t in sa + so:
            t.localize(s0)
        for t in a_s:
            t.localize(self.ae[0])
        for t in o_s:
            t.localize(self.oe[0])
        self.sa_cons, self.sa_bil = sa
        self.so_cons, self.so_bil = so if so else (None, None)
        self.as_bil, self.as_cons = a_s
        self.os_bil, self.os_cons = o_s
        self.nA_ext = (self.ae[1] - self.ae[0]) * A.im
        self.nO_ext = (self.oe[1] - self.oe[0]) * O.im
        self.nS = (s1 - s0) * S.im
        self.nA, self.nO = (a1 - a0) * A.im, (o1 - o0) * O.im
        self.own_off = (a0 - self.ae[0]) * A.im            # owned atmosphere columns inside the extended band
        # the atmosphere: column blocks = MPI ranks, one module instance (matrices) each
        self.ranks = max(1, min(ranks, self.nA_ext // 1024 or 1))
        cut = [self.nA_ext * r // self.ranks for r in range(self.ranks + 1)]
        self.cut = list(zip(cut[:-1], cut[1:]))
        self.vd = [orc.VDiff(b - a, 1, K, nc, iq, consts["Grav"], consts["CpDry"], consts["GasRDry"], consts["DelTime"])
                   for a, b in self.cut]
        # result arrays exist before the calls, as the reference's module arrays do
        self.fwd_out = [{"DUDt": np.zeros((K, b - a)), "DVDt": np.zeros((K, b - a)), "DTempDt": np.zeros((K, b - a)),
                         "DQMixDt": np.zeros((nc, K, b - a)), "ImplCplCoef1": np.zeros((4, b - a)),
                         "ImplCplCoef2": np.zeros((4, b - a))} for a, b in self.cut]
        self.bulk_out = None
        self.parallel_atm = True
        self._col_blocks = None
        self._pool = None

    # ------------------------------------------------------------------ what the band needs as input
    def input_rows(self):
        """(atmosphere rows, ocean rows) whose fields the band reads: the owned rows plus the remaps' halo rows"""
        return self.ae, self.oe

    def fraction(self):
        """share of the whole exchange this band's work is (extended atmosphere rows / all rows)"""
        return (self.ae[1] - self.ae[0]) / self.A.jm

    def set_inputs(self, col, atm, ocn):
        """col: VDiffForward inputs on the extended atmosphere rows (dict, (levels, nA_ext) / (nc, levels, nA_ext));
        atm: WindU .. SnowFall on the same rows (nA_ext,); ocn: SfcTempO/I, SfcAlbedoO/I, SIceCon (nO_ext,)"""
        self._col_blocks = [{k: np.ascontiguousarray(v[..., a:b], dtype=np.float64) for k, v in col.items()} for a, b in self.cut]
        self.atm = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in atm.items()}
        self.ocn = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in ocn.items()}
        self.a2s_bil = np.empty((13, self.nA_ext))
        for l, k in enumerate(("WindU", "WindV", "SfcAirTemp", "QVap1", "SfcPress")):
            self.a2s_bil[l] = self.atm[k]
        self.a2s_cons = np.stack([self.atm[k] for k in ("LDwRFlx", "SDwRFlx", "RainFall", "SnowFall")])
        self.o2s_bil = np.stack([self.ocn["SfcTempO"], self.ocn["SfcTempI"]])
        self.o2s_cons = np.stack([self.ocn["SIceCon"], self.ocn["SfcAlbedoO"], self.ocn["SfcAlbedoI"]])
        S = self.S
        JA, IA = self.s_rows[1] - self.s_rows[0] + 2, S.im + 2
        one = lambda fill=1.0: np.full((JA, IA), fill)
        self.halo = {"WindU": one(), "WindV": one(), "SfcAirTemp": one(280.0), "QVap1": one(), "SfcPress": one(1e5),
                     "ImplCplCoef1": np.ones((4, JA, IA)), "ImplCplCoef2": np.ones((4, JA, IA)),
                     "LDwRFlx": one(), "SDwRFlx": one(),
                     "SfcTemp": np.stack([one(280.0), one(270.0), one(0.0)]),
                     "SfcAlbedo": np.stack([one(), one(), one(0.0)]),
                     "SIceCon": one(0.0), "SfcHeight": one(0.0), "Sig1Info": np.array([self.c["Sig1"], 0.01])}
        self.s_bil, self.s_cons = np.zeros((13, self.nS)), np.zeros((4, self.nS))
        self.s_obil, self.s_ocons = np.zeros((2, self.nS)), np.zeros((3, self.nS))
        self.s2a, self.s2o = np.zeros((9, self.nS)), np.zeros((12, self.nS))
        self.a_recv, self.o_recv = np.zeros((9, self.nA)), np.zeros((12, self.nO))

    # ------------------------------------------------------------------ one exchange
    def _atm(self, fn):
        if not self.parallel_atm or self.ranks == 1:
            return [fn(r) for r in range(self.ranks)]

        def work(r):
            self.orc.set_num_threads(1)          # a rank's own OpenMP regions: one thread, the rank IS the parallelism
            return fn(r)
        if self._pool is None:                   # the rank threads live as long as the band: no thread start-up in the timed calls
            self._pool = ThreadPoolExecutor(max_workers=self.ranks)
        return list(self._pool.map(work, range(self.ranks)))

    def _remap(self, t, x, y):
        self.orc.remap_apply(t.send, t.recv, t.coef, x, t.n_dst, recv=y)

    def run(self):
        """one exchange; returns (seconds, per-stage seconds).  Results: self.a_recv, self.o_recv, self.tend (owned rows)."""
        o, S = self.orc, self.S
        t0 = time.perf_counter()
        f = self._atm(lambda r: self.vd[r].forward(self._col_blocks[r], out=self.fwd_out[r]))
        for r, (a, b) in enumerate(self.cut):                      # put side of the atmosphere (ref atm/dccm_atm_mod.f90:711-712)
            self.a2s_bil[5:9, a:b] = f[r]["ImplCplCoef1"]
            self.a2s_bil[9:13, a:b] = f[r]["ImplCplCoef2"]
        t1 = time.perf_counter()
        self._remap(self.as_bil, self.a2s_bil, self.s_bil)
        self._remap(self.as_cons, self.a2s_cons, self.s_cons)
        self._remap(self.os_bil, self.o2s_bil, self.s_obil)
        self._remap(self.os_cons, self.o2s_cons, self.s_ocons)
        t2 = time.perf_counter()
        h = self.halo
        js = self.s_rows[1] - self.s_rows[0]
        I = (slice(1, -1), slice(1, -1))
        put = lambda dst, src: dst[I].__setitem__(slice(None), src.reshape(js, S.im))     # unpack, ref sfc/dccm_sfc_mod.f90:900-951
        for l, k in enumerate(("WindU", "WindV", "SfcAirTemp", "QVap1", "SfcPress")):
            put(h[k], self.s_bil[l])
        for c in range(4):
            put(h["ImplCplCoef1"][c], self.s_bil[5 + c]); put(h["ImplCplCoef2"][c], self.s_bil[9 + c])
        put(h["LDwRFlx"], self.s_cons[0]); put(h["SDwRFlx"], self.s_cons[1])
        put(h["SfcTemp"][0], self.s_obil[0]); put(h["SfcTemp"][1], self.s_obil[1])
        put(h["SIceCon"], self.s_ocons[0]); put(h["SfcAlbedo"][0], self.s_ocons[1]); put(h["SfcAlbedo"][1], self.s_ocons[2])
        b = self.bulk_out = o.bulkflux(S.im + 2, js + 2, h, out=self.bulk_out)
        g = lambda k, n: b[k][n][I].reshape(-1)                                           # pack, ref :787-809
        s2a, s2o = self.s2a, self.s2o
        s2a[0] = g("LUwRFlx", 2); s2a[1] = g("SUwRFlx", 2); s2a[2] = g("SenHFlx", 2); s2a[3] = g("QVapMFlx", 2)
        s2a[4] = g("SfcAlbedo", 2)
        for c in range(4):
            s2a[5 + c] = g("DelVarImplCPL", c)
        s2o[0] = g("SfcHFlx_ns", 0); s2o[1] = g("SfcHFlx_sr", 0); s2o[2] = self.s_cons[3]; s2o[3] = self.s_cons[2]
        s2o[4] = g("QVapMFlx", 0); s2o[5] = -g("WindStressX", 2); s2o[6] = -g("WindStressY", 2)
        s2o[7] = g("SfcHFlx_ns", 1); s2o[8] = g("SfcHFlx_sr", 1); s2o[9] = g("QVapMFlx", 1)
        s2o[10] = g("DSfcHFlxDTs", 0); s2o[11] = g("DSfcHFlxDTs", 1)
        t3 = time.perf_counter()
        self._remap(self.sa_cons, s2a[:4], self.a_recv[:4])
        self._remap(self.sa_bil, s2a[4:], self.a_recv[4:])
        if self.so_cons is not None:
            self._remap(self.so_cons, s2o[:10], self.o_recv[:10])
            self._remap(self.so_bil, s2o[10:], self.o_recv[10:])
        t4 = time.perf_counter()
        off, iq = self.own_off, self.iq
        for r, (a, bb) in enumerate(self.cut):                      # level-1 values from the surface, ref atm/dccm_atm_mod.f90:832-835
            lo, hi = max(a, off), min(bb, off + self.nA)
            if lo < hi:
                src = slice(lo - off, hi - off)
                f[r]["DUDt"][0, lo - a:hi - a] = self.a_recv[5, src]
                f[r]["DVDt"][0, lo - a:hi - a] = self.a_recv[6, src]
                f[r]["DTempDt"][0, lo - a:hi - a] = self.a_recv[7, src]
                f[r]["DQMixDt"][iq - 1, 0, lo - a:hi - a] = self.a_recv[8, src]
        self._atm(lambda r: self.vd[r].backward(f[r]["DUDt"], f[r]["DVDt"], f[r]["DTempDt"], f[r]["DQMixDt"], inplace=True))
        t5 = time.perf_counter()
        self._f = f
        return t5 - t0, {"fwd": t1 - t0, "remap_to_sfc": t2 - t1, "bulk": t3 - t2, "remap_from_sfc": t4 - t3, "bwd": t5 - t4}

    @property
    def tend(self):
        """the four tendencies on the OWNED atmosphere rows"""
        off = self.own_off
        cat = lambda k: np.concatenate([f[k] for f in self._f], axis=-1)[..., off:off + self.nA]
        return {k: cat(k) for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt")}

    def coef(self):
        """ImplCplCoef1 / 2 on the owned rows"""
        off = self.own_off
        return self.a2s_bil[5:9, off:off + self.nA], self.a2s_bil[9:13, off:off + self.nA]
