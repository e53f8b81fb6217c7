"""Generates tests/golden/*.npz FROM THE ORACLE (python tests/golden/make_golden.py).

The reference cannot be executed in this environment (Fortran-only, no Fortran compiler, and
it ships no fixtures), so these goldens do not pin the oracle to the reference -- they pin the
oracle (and through it the CUDA path) against accidental drift, at sizes small enough to
commit.  Inputs are regenerated from dennou-ccm_b200/synthetic.py (pure functions of the cell
index), so only outputs are stored."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def bulk_case(S, g):
    from test_oracle_kat import _bulk_inputs
    return _bulk_inputs(S, g)


def compute(orc, dccm, S):
    from util import as_orc_grid, pair
    out = {}
    A, O, Sx = [as_orc_grid(orc, g) for g in pair(orc, dccm, "T21_Pl42")]
    t = {}
    for name, tab in (("as2", orc.gen_jones99(A, Sx, 2)), ("so1", orc.gen_jones99(Sx, O, 1)),
                      ("os1", orc.gen_jones99(O, Sx, 1)), ("as_bil", orc.gen_bilinear(A, Sx)),
                      ("so_bil", orc.gen_bilinear(Sx, O))):
        send, recv, coef = tab.to_index(*((A.im, Sx.im) if name.startswith("as") else
                                          (Sx.im, O.im) if name.startswith("so") else (O.im, Sx.im)))
        t[name + "_send"], t[name + "_recv"], t[name + "_coef"] = send, recv, coef
    out["tables_T21_Pl42"] = t

    g = dccm.tables.get_LonLatGrid(16, 8)
    IA, JA, inp = bulk_case(S, g)
    b = orc.bulkflux(IA, JA, inp)
    out["bulkflux_16x8"] = {k: np.ascontiguousarray(v[:, 1:-1, 1:-1]) for k, v in b.items()}

    g = dccm.tables.get_LonLatGrid(8, 4)
    K, nc = 6, 2
    vin = S.column_inputs(np, g, K, nc)
    vd = orc.VDiff(g.im, g.jm, K, nc, 2, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    f = vd.forward(vin)
    lvl1 = 1e-3 * np.stack([S.normal(np, np.arange(g.n, dtype=np.float64), 70.0 + k) for k in range(4)])
    DU, DV, DT, DQ = f["DUDt"].copy(), f["DVDt"].copy(), f["DTempDt"].copy(), f["DQMixDt"].copy()
    DU[0], DV[0], DT[0], DQ[1, 0] = lvl1
    bw = vd.backward(DU, DV, DT, DQ)
    v = {"fwd_" + k: a for k, a in f.items()}
    v.update(bwd_DUDt=bw[0], bwd_DVDt=bw[1], bwd_DTempDt=bw[2], bwd_DQMixDt=bw[3],
             diag_UV=vd.diag(0)[1:], diag_T=vd.diag(1)[1:], diag_Q=vd.diag(2)[1:])
    out["vdiff_8x4_K6"] = v
    return out


if __name__ == "__main__":
    import oracle
    oracle.build()
    dccm = importlib.import_module("dennou-ccm_b200")
    S = importlib.import_module("dennou-ccm_b200.synthetic")
    for name, arrs in compute(oracle, dccm, S).items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, {k: v.shape for k, v in arrs.items()})
