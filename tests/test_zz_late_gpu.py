"""Second-order conservative A->S table through the whole exchange (run with -m gpu on a B200).

gmapgen's default for the atmosphere -> exchange-grid table is interp_order_AS = 2 (ref tool/gmapgen/gmapgen_main.f90:219),
and the reference's shipped conservative configuration (exp/APEI07Couple/common/genmapgen_ATM_T42-OCN_Pl42_conserve.conf)
does not override it: every destination row then carries its first-order entry plus the -w2/dphi, +w2/dphi pair on the
two neighbouring source latitudes (ref common/grid_mapping_util_jones99.f90:252-267) -- three source rows per stencil
in the fused surface kernel's tiles instead of one.
"""
import importlib

import numpy as np
import pytest

from util import pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S(dccm):
    return importlib.import_module("dennou-ccm_b200.synthetic")


@pytest.mark.parametrize("name,members", [("T21_Pl42", 1), ("T21_1deg", 2), ("T106_1deg", 1)])
def test_fused_surface_kernel_with_second_order_atmosphere_table(gpu, orc, dccm, S, name, members):
    """staged and direct forms of the fused kernel == unfused remap -> bulk flux -> pack, bit for bit"""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    L = dccm._lib
    A, O, Sx = pair(orc, dccm, name)
    tabs = X.build_tables(A, O, Sx, order_as=2)
    assert len(tabs["as_cons"][2]) > 2 * Sx.n                     # the pairs are there
    M, K = members, 8
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    cols = [tt(S.column_inputs(np, A, K, 1, member=m)) for m in range(M)]
    atms = [tt(S.atm_surface_fields(np, A, member=m)) for m in range(M)]
    ocns = [tt(S.ocn_surface_fields(np, O, member=m)) for m in range(M)]
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, members=M, device=gpu)
    ex.set_inputs({k: torch.cat([c[k] for c in cols], dim=-1).contiguous() for k in cols[0]},
                  {k: torch.stack([a[k] for a in atms]) for k in atms[0]},
                  {k: torch.stack([o[k] for o in ocns]) for k in ocns[0]})
    ex.forward()
    ex.remap_to_sfc(); ex.bulk(); ex.pack_sfc()
    torch.cuda.synchronize()
    want = {"s2a": ex.s2a.clone(), "s2o": ex.s2o.clone()}
    forms = []
    try:
        for staged in (1, 0):
            ex.configure_sfc(staged, 5)
            ex.s2a.fill_(float("nan")); ex.s2o.fill_(float("nan"))
            ex.sfc_fused()
            torch.cuda.synchronize()
            forms.append(ex.sfc_last_form())
            assert torch.equal(ex.s2a, want["s2a"]), f"staged={staged} s2a"
            assert torch.equal(ex.s2o, want["s2o"]), f"staged={staged} s2o"
    finally:
        ex.configure_sfc(1, 5)
    print(name, "forms launched (1 = staged):", forms)
    assert forms[1] == 0


def test_exchange_step_vs_oracle_second_order(gpu, orc, dccm, S):
    """whole exchange with the second-order A->S table against the oracle: remaps and the reference-order column
    solve bit-exact, bulk-flux-derived stages within the conditioning bar"""
    import torch
    from exchange_ref import compare_exchange, oracle_exchange
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, "T21_Pl42")
    K = 16
    tabs = X.build_tables(A, O, Sx, order_as=2)
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, fast=False, device=gpu)
    col, atm, ocn = S.column_inputs(np, A, K, 1), S.atm_surface_fields(np, A), S.ocn_surface_fields(np, O)
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    ex.set_inputs(tt(col), {k: v[None] for k, v in tt(atm).items()}, {k: v[None] for k, v in tt(ocn).items()})
    ex.step(fused=False)
    torch.cuda.synchronize()
    ref = oracle_exchange(orc, S, A, O, Sx, K, 1, 1, tabs, col, atm, ocn)
    detail = {}
    worst = compare_exchange(ex, ref, detail=detail)
    exact = ("Coef1", "Coef2", "s_bil", "s_cons", "s_obil", "s_ocons")
    assert all(detail[k] == 0.0 for k in exact), {k: detail[k] for k in exact}
    assert worst <= 1e-11, detail
    keep = {k: getattr(ex, k).clone() for k in ("s2a", "s2o", "a_recv", "o_recv")}
    ex.step(fused=True)
    torch.cuda.synchronize()
    for k, v in keep.items():
        assert torch.equal(getattr(ex, k), v), f"fused step differs in {k}"


def test_atm_get_side_assembly(gpu, orc, dccm, S):
    """ref atm/dccm_atm_mod.f90:823-836 on the device: radiative surface temperature from the remapped composite LUwRFlx
    (device pow vs libm pow: within 1e-15), the other gets are exact copies of their layers; cells beyond n untouched."""
    import torch
    A = dccm.tables.get_LonLatGrid(128, 64)
    n, ld = A.n, A.n + 32
    rng = np.random.default_rng(5)
    r = rng.normal(0.0, 50.0, (9, ld))
    r[0] = 5.670373e-8 * (240.0 + 70.0 * rng.random(ld)) ** 4
    a_recv = torch.as_tensor(r, device=gpu).contiguous()
    got = dccm.dccm_atm_mod.atm_get_assemble(a_recv, n)
    want = orc.atm_sfc_temp(r[0, :n])
    np.testing.assert_allclose(got["SfcTemp"].cpu().numpy(), want, rtol=1e-15, atol=0)     # CUDA pow: 2 ulp, glibc: < 1 ulp
    assert np.array_equal(got["SfcAlbedo"].cpu().numpy(), r[4, :n])
    assert np.array_equal(got["SurfHeatFlux"].cpu().numpy(), r[2, :n])
    assert np.array_equal(got["SurfH2OVapFlux"].cpu().numpy(), r[3, :n])
    with pytest.raises(dccm.DccmError, match="n <= ld"):
        dccm.dccm_atm_mod.atm_get_assemble(a_recv, ld + 1)


def test_atm_legacy_get_side_assembly(gpu, orc, dccm):
    """legacy 2-component mode (ref atm/mod_atm.f90:740-775, atm/dcpam_main_mod.f90:1003-1031) on the device: fourth
    root within 1e-15 (device pow vs libm), everything else the oracle's bits; cells beyond n untouched."""
    import torch
    n, ld = 64 * 32 + 5, 64 * 32 + 16
    rng = np.random.default_rng(9)
    r = np.empty((4, ld))
    r[0] = (250.0 + 50.0 * rng.random(ld)) ** 4; r[1] = 0.05 + 0.6 * rng.random(ld)
    r[2] = rng.normal(0.0, 3.0, ld); r[3] = rng.random(ld) * 0.4
    p0 = 1.0e5 + rng.normal(0.0, 500.0, n); p1 = p0 - (900.0 + 100.0 * rng.random(n))
    tb = 280.0 + rng.normal(0.0, 5.0, n)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=gpu)
    tb_dev = t(tb)
    got = dccm.dccm_atm_mod.atm_legacy_get_assemble(t(r), 14400.0, 9.8, 1004.6, t(p0), t(p1), tb_dev, n)
    st, sn, tb_want = orc.atm_legacy_get(r[0, :n], r[3, :n], r[2, :n], 14400.0, 9.8, 1004.6, p0, p1, tb)
    np.testing.assert_allclose(got["SurfTemp"].cpu().numpy(), st, rtol=1e-15, atol=0)
    assert np.array_equal(got["SurfSnow"].cpu().numpy(), sn)
    assert np.array_equal(got["SurfAlbedo"].cpu().numpy(), r[1, :n])
    assert np.array_equal(tb_dev.cpu().numpy(), tb_want)



@pytest.mark.parametrize("name,nslab", [("T21_1deg", 4), ("T106_1deg", 7), ("T42_T42", 16)])
def test_slab_pipelined_step_equals_the_plain_step(gpu, orc, dccm, S, name, nslab):
    """SurfaceExchange.step_pipelined (forward solve in latitude slabs, the fused surface kernel of every slab on a
    second stream as soon as the rows it reads are solved; row-range launches of the same kernels) gives the bits of
    step(), eagerly and replayed from a CUDA graph."""
    import torch
    from util import pair
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, name)
    K = 26
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, fast=False, device=gpu)
    col, atm, ocn = S.column_inputs(torch, A, K, 1, dev=gpu), S.atm_surface_fields(torch, A, dev=gpu), S.ocn_surface_fields(torch, O, dev=gpu)
    ex.set_inputs(col, {k: v[None] for k, v in atm.items()}, {k: v[None] for k, v in ocn.items()})
    ex.step()
    torch.cuda.synchronize()
    names = ("s2a", "s2o", "a_recv", "o_recv")
    keep = {k: getattr(ex, k).clone() for k in names}
    keep.update({k: v.clone() for k, v in ex.tend.items()})

    def check(tag):
        torch.cuda.synchronize()
        for k in names:
            assert torch.equal(getattr(ex, k), keep[k]), (tag, k)
        for k, v in ex.tend.items():
            assert torch.equal(v, keep[k]), (tag, k)

    def scrub():
        for k in names:
            getattr(ex, k).fill_(float("nan"))
        for v in ex.tend.values():
            v.fill_(float("nan"))
    scrub(); ex.step_pipelined(nslab); check("eager")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ex.step_pipelined(nslab)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ex.step_pipelined(nslab)
    scrub(); g.replay(); check("graph")


@pytest.mark.parametrize("name,K,nc,iq", [("T21_1deg", 9, 3, 2), ("T21_Pl42", 12, 2, 1)])
def test_exchange_with_several_tracers_vs_oracle(gpu, orc, dccm, S, name, K, nc, iq):
    """Whole exchange with ncmax > 1 and the water-vapour tracer not first (composition: IndexH2OVap, ref
    atm/dcpam_sfc_implicit_coupling_mod.f90:316, :371-374; the level-1 update goes to tracer IndexH2OVap only,
    atm/dccm_atm_mod.f90:835): every stage output has the oracle's bits, unfused and fused alike."""
    import torch
    from exchange_ref import compare_exchange, oracle_exchange
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, name)
    tabs = X.build_tables(A, O, Sx)
    ex = X.SurfaceExchange(A, O, Sx, K, nc, iq, tabs=tabs, fast=False, device=gpu)
    col, atm, ocn = S.column_inputs(np, A, K, nc), S.atm_surface_fields(np, A), S.ocn_surface_fields(np, O)
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    ex.set_inputs(tt(col), {k: v[None] for k, v in tt(atm).items()}, {k: v[None] for k, v in tt(ocn).items()})
    ref = oracle_exchange(orc, S, A, O, Sx, K, nc, iq, tabs, col, atm, ocn)
    for fused in (False, True):
        for t in (ex.s2a, ex.s2o, ex.a_recv, ex.o_recv, *ex.tend.values()):
            t.fill_(float("nan"))
        ex.step(fused=fused)
        torch.cuda.synchronize()
        same = {}
        compare_exchange(ex, ref, bitwise=same)
        if fused:                                   # the fused kernel does not store the remapped surface inputs
            same = {k: v for k, v in same.items() if not k.startswith("s_")}
        assert all(same.values()), (fused, same)
    # the other tracers' level 1 is NOT touched by the surface update
    other = [n for n in range(nc) if n != iq - 1]
    lvl1 = ref["fwd"]["DQMixDt"][other, 0] / (2.0 * S.DELTIME)
    assert np.array_equal(ex.tend["DQMixDt"][other, 0].cpu().numpy(), lvl1)
