"""The whole exchange step composed from ORACLE calls (checker for the device-resident
SurfaceExchange): forward -> remap A->S, O/I->S -> bulk flux -> remap S->A, S->O/I -> backward,
with the field lists of the reference glue."""
import numpy as np


def _halo(x, jm, im, fill=1.0):
    full = np.full((jm + 2, im + 2), fill)
    full[1:-1, 1:-1] = x.reshape(jm, im)
    return full


def oracle_exchange(orc, syn, A, O, S, K, nc, iq, tabs, col, atm, ocn, consts=None):
    c = consts or {}
    Grav, Cp, R, dt = c.get("Grav", syn.GRAV), c.get("CpDry", syn.CPDRY), c.get("GasRDry", syn.GASRDRY), c.get("DelTime", syn.DELTIME)
    sig1 = c.get("Sig1", syn.SIG1)
    vd = orc.VDiff(A.im, A.jm, K, nc, iq, Grav, Cp, R, dt)
    f = vd.forward(col)
    r = {"fwd": f}
    a2s_bil = np.stack([atm["WindU"], atm["WindV"], atm["SfcAirTemp"], atm["QVap1"], atm["SfcPress"],
                        *f["ImplCplCoef1"], *f["ImplCplCoef2"]])
    a2s_cons = np.stack([atm[k] for k in ("LDwRFlx", "SDwRFlx", "RainFall", "SnowFall")])
    o2s_bil = np.stack([ocn["SfcTempO"], ocn["SfcTempI"]])
    o2s_cons = np.stack([ocn["SIceCon"], ocn["SfcAlbedoO"], ocn["SfcAlbedoI"]])
    rm = lambda key, x, n: orc.remap_apply(*tabs[key], x, n)
    s_bil, s_cons = rm("as_bil", a2s_bil, S.n), rm("as_cons", a2s_cons, S.n)
    s_obil, s_ocons = rm("os_bil", o2s_bil, S.n), rm("os_cons", o2s_cons, S.n)
    r.update(s_bil=s_bil, s_cons=s_cons, s_obil=s_obil, s_ocons=s_ocons)
    H = lambda x, fill=1.0: _halo(x, S.jm, S.im, fill)
    inp = {"WindU": H(s_bil[0]), "WindV": H(s_bil[1]), "SfcAirTemp": H(s_bil[2], 280.0), "QVap1": H(s_bil[3]),
           "SfcPress": H(s_bil[4], 1e5), "ImplCplCoef1": np.stack([H(x) for x in s_bil[5:9]]),
           "ImplCplCoef2": np.stack([H(x) for x in s_bil[9:13]]),
           "LDwRFlx": H(s_cons[0]), "SDwRFlx": H(s_cons[1]),
           "SfcTemp": np.stack([H(s_obil[0], 280.0), H(s_obil[1], 270.0), H(np.zeros(S.n))]),
           "SfcAlbedo": np.stack([H(s_ocons[1]), H(s_ocons[2]), H(np.zeros(S.n))]),
           "SIceCon": H(s_ocons[0], 0.0), "SfcHeight": np.zeros((S.jm + 2, S.im + 2)),
           "Sig1Info": np.array([sig1, 0.01])}
    b = orc.bulkflux(S.im + 2, S.jm + 2, inp)
    b = {k: np.ascontiguousarray(v[:, 1:-1, 1:-1]).reshape(v.shape[0], S.n) for k, v in b.items()}
    r["bulk"] = b
    s2a = np.stack([b["LUwRFlx"][2], b["SUwRFlx"][2], b["SenHFlx"][2], b["QVapMFlx"][2], b["SfcAlbedo"][2],
                    *b["DelVarImplCPL"]])
    s2o = np.stack([b["SfcHFlx_ns"][0], b["SfcHFlx_sr"][0], s_cons[3], s_cons[2], b["QVapMFlx"][0],
                    -b["WindStressX"][2], -b["WindStressY"][2],
                    b["SfcHFlx_ns"][1], b["SfcHFlx_sr"][1], b["QVapMFlx"][1],
                    b["DSfcHFlxDTs"][0], b["DSfcHFlxDTs"][1]])
    a_recv = np.concatenate([rm("sa_cons", s2a[:4], A.n), rm("sa_bil", s2a[4:], A.n)])
    o_recv = np.concatenate([rm("so_cons", s2o[:10], O.n), rm("so_bil", s2o[10:], O.n)])
    r.update(s2a=s2a, s2o=s2o, a_recv=a_recv, o_recv=o_recv)
    DU, DV, DT, DQ = f["DUDt"].copy(), f["DVDt"].copy(), f["DTempDt"].copy(), f["DQMixDt"].copy()
    DU[0], DV[0], DT[0], DQ[iq - 1, 0] = a_recv[5:9]
    r["bwd"] = dict(zip(("DUDt", "DVDt", "DTempDt", "DQMixDt"), vd.backward(DU, DV, DT, DQ)))
    return r


def floor_rel(a, b, frac=1e-3):
    """per-cell |a-b| / max(|b|, frac * max|b| of the layer): flux sums cancel, so the denominator
    is floored at 0.1 % of the layer's magnitude ('tolerances')."""
    a, b = np.asarray(a), np.asarray(b)
    a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    worst = 0.0
    for x, y in zip(a2, b2):
        if np.any(np.isnan(x) != np.isnan(y)):
            return float("inf")
        m = ~np.isnan(y)
        if not m.any():
            continue
        den = np.maximum(np.abs(y[m]), frac * max(np.abs(y[m]).max(), 1e-300))
        worst = max(worst, float((np.abs(x[m] - y[m]) / den).max()))
    return worst


def bits_equal(a, b):
    """same shape and the same 64-bit patterns in every cell (so +0 / -0 and NaN payloads count too)."""
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return a.size == b.size and np.array_equal(a.reshape(-1).view(np.int64), b.reshape(-1).view(np.int64))


def compare_exchange(ex, ref, members=1, detail=None, bitwise=None):
    """max floored relative error over every stage output of a 1-member SurfaceExchange; `bitwise` (a dict)
    receives, per stage, whether the device array and the oracle's have identical bits."""
    g = lambda t: t.detach().cpu().numpy()
    M = ex.M
    checks = [("Coef1", g(ex.a2s_bil[5 * M:9 * M]), ref["fwd"]["ImplCplCoef1"]),
              ("Coef2", g(ex.a2s_bil[9 * M:13 * M]), ref["fwd"]["ImplCplCoef2"]),
              ("s_bil", g(ex.s_bil), ref["s_bil"]), ("s_cons", g(ex.s_cons), ref["s_cons"]),
              ("s_obil", g(ex.s_obil[:2 * M]), ref["s_obil"]), ("s_ocons", g(ex.s_ocons[:3 * M]), ref["s_ocons"]),
              ("s2a", g(ex.s2a), ref["s2a"]), ("s2o", g(ex.s2o), ref["s2o"]),
              ("a_recv", g(ex.a_recv), ref["a_recv"]), ("o_recv", g(ex.o_recv), ref["o_recv"])]
    for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"):
        checks.append(("bwd_" + k, g(ex.tend[k]), ref["bwd"][k]))
    worst = 0.0
    for name, a, b in checks:
        e = floor_rel(a, b)
        if bitwise is not None:
            bitwise[name] = bits_equal(a, b)
        if detail is not None:
            detail[name] = e
        worst = max(worst, e)
    return worst
