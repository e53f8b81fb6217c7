"""oracle/band.py (the reference's exchange on a latitude band, oracle generators and oracle calls only -- what bench.py
times as the CPU baseline and checks the GPU against) versus tests/exchange_ref.oracle_exchange (whole grids, tables
from the product's generators): same bits on the whole grid, and a band reproduces exactly its rows of the whole."""
import importlib

import numpy as np
import pytest

from exchange_ref import bits_equal, oracle_exchange
from util import pair


def _whole(orc, dccm, S, name, K, order):
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, name)
    tabs = X.build_tables(A, O, Sx, order_as=order)
    col, atm, ocn = S.column_inputs(np, A, K, 1), S.atm_surface_fields(np, A), S.ocn_surface_fields(np, O)
    return (A, O, Sx), (col, atm, ocn), oracle_exchange(orc, S, A, O, Sx, K, 1, 1, tabs, col, atm, ocn)


CONSTS = lambda S: {"Grav": S.GRAV, "CpDry": S.CPDRY, "GasRDry": S.GASRDRY, "DelTime": S.DELTIME, "Sig1": S.SIG1}


@pytest.mark.parametrize("name,K,order,ranks", [("T21_Pl42", 16, 1, 1), ("T21_1deg", 26, 1, 3), ("T42_T42", 26, 2, 2)])
def test_band_exchange_whole_grid_and_bands(orc, dccm, name, K, order, ranks):
    from oracle.band import BandExchange, make_grids
    S = importlib.import_module("dennou-ccm_b200.synthetic")
    (A, O, Sx), (col, atm, ocn), ref = _whole(orc, dccm, S, name, K, order)
    # the oracle's own grids are the product's grids, bit for bit
    A2, O2, S2 = make_grids(orc, A.im, A.jm, O.im, O.jm, name.endswith("1deg"))
    for g, h in ((A, A2), (O, O2), (Sx, S2)):
        assert g.im == h.im and g.jm == h.jm
        for k in ("x_Lon", "y_Lat", "x_LonWt", "y_LatWt"):
            assert bits_equal(getattr(g, k), getattr(h, k)), k
    for rows in ((0, A.jm), (A.jm // 2 - 3, A.jm // 2 + 5), (0, 4), (A.jm - 5, A.jm), (7, 8)):
        bx = BandExchange(orc, A2, O2, S2, K, 1, rows, CONSTS(S), order_as=order, ranks=ranks)
        (ae0, ae1), (oe0, oe1) = bx.input_rows()
        assert ae0 <= rows[0] and ae1 >= rows[1] and ae1 - ae0 <= rows[1] - rows[0] + 6
        cut = lambda d, j0, j1, im: {k: v[..., j0 * im:j1 * im] for k, v in d.items()}
        bx.set_inputs(cut(col, ae0, ae1, A.im), cut(atm, ae0, ae1, A.im), cut(ocn, oe0, oe1, O.im))
        bx.run()
        a0, a1 = rows
        o0, o1 = bx.o_rows
        assert bits_equal(bx.a_recv, ref["a_recv"][:, a0 * A.im:a1 * A.im]), rows
        assert bits_equal(bx.o_recv, ref["o_recv"][:, o0 * O.im:o1 * O.im]), rows
        s0, s1 = bx.s_rows
        assert bits_equal(bx.s2a, ref["s2a"][:, s0 * Sx.im:s1 * Sx.im]) and bits_equal(bx.s2o, ref["s2o"][:, s0 * Sx.im:s1 * Sx.im])
        for k, v in bx.tend.items():
            assert bits_equal(v, ref["bwd"][k][..., a0 * A.im:a1 * A.im]), (rows, k)
        c1, c2 = bx.coef()
        assert bits_equal(c1, ref["fwd"]["ImplCplCoef1"][:, a0 * A.im:a1 * A.im])
        if rows == (0, A.jm):
            assert bx.o_rows == (0, O.jm) and bx.s_rows == (0, Sx.jm) and bx.fraction() == 1.0
