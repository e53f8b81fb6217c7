"""Parity of the CUDA path against the CPU oracle, through the C ABI (run with -m gpu on a B200).

Bars (BASELINE.json north_star): identical remap indices; per-cell fp64 relative error
<= 1e-12; global-integral conservation <= 1e-13 relative.  Where the kernel keeps the
reference's operation order (remap, column solves in reference-order mode, bulk flux) the comparison is
BIT-EXACT: the bulk flux's exp / log / ** are fixed IEEE sequences on both sides (csrc/dccm_pmath.cuh,
oracle/orc_pmath.h), so no tolerance or floor is needed anywhere on the reference-order path.
"""
import importlib
import os

import numpy as np
import pytest

from util import as_orc_grid, pair, relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-12          # north_star: <= 1e-12 relative per cell in fp64
CONS = 1e-13          # north_star: global-integral conservation <= 1e-13 relative


@pytest.fixture(scope="module")
def S(dccm):
    return importlib.import_module("dennou-ccm_b200.synthetic")


def _tables(orc, dccm, name):
    T = dccm.tables
    A, O, Sx = pair(orc, dccm, name)
    out = []
    for label, s, d in (("A->S", A, Sx), ("S->A", Sx, A), ("S->O", Sx, O), ("O->S", O, Sx)):
        out.append((label + " cons1", s, d, T.gen_table_jones99(s, d, 1, 1)))
        out.append((label + " bilin", s, d, T.gen_table_bilinear(s, d, 1)))
    try:
        out.append(("A->S cons2", A, Sx, T.gen_table_jones99(A, Sx, 2, 1)))
    except dccm.DccmError:
        pass
    return out


# ------------------------------------------------------------------ K1 remap

@pytest.mark.parametrize("name", ["T21_Pl42", "T42_T42", "T106_1deg"])
def test_remap_bit_exact_vs_oracle(gpu, orc, dccm, S, name):
    for label, s, d, tab in _tables(orc, dccm, name):
        send_i, recv_i, coef = tab.index(s.im, d.im)
        op = dccm.RemapOperator(send_i, recv_i, coef, s.n, d.n)
        assert op.nnz == len(coef)
        x = S.generic_fields(np, s, 16)
        got = op.apply_host(x)
        ref = orc.remap_apply(send_i, recv_i, coef, x, d.n)
        assert np.array_equal(got, ref), f"{name} {label}: max rel err {relerr(got, ref)}"
        # global integral (area-weighted) agrees to the conservation bar (here: exactly)
        w = np.repeat(d.y_LatWt, d.im)
        assert np.abs((got * w).sum(1) / (ref * w).sum(1) - 1.0).max() <= CONS


@pytest.mark.parametrize("name", ["T21_Pl42", "T42_T42", "T106_1deg"])
def test_remap_zonal_stencil_form_bit_exact(gpu, orc, dccm, S, name):
    """Given the grids' row lengths the library stores zonally repeating tables as one stencil per
    latitude row (kind 1); results are the same bits as the CSR form and the oracle."""
    kinds = []
    for label, s, d, tab in _tables(orc, dccm, name):
        send_i, recv_i, coef = tab.index(s.im, d.im)
        op = dccm.RemapOperator(send_i, recv_i, coef, s.n, d.n, gnxs=s.im, gnxr=d.im)
        kinds.append(op.kind)
        x = S.generic_fields(np, s, 11)
        got = op.apply_host(x, rn2=13, num_of_data=11)
        ref = orc.remap_apply(send_i, recv_i, coef, x, d.n, 13, 11)
        assert np.array_equal(got, ref), f"{name} {label} kind {op.kind}"
    assert 1 in kinds, kinds


@pytest.mark.parametrize("name", ["T21_1deg", "T106_1deg"])
def test_remap_separable_form_bit_exact(gpu, orc, dccm, S, name):
    """operators created straight from the grids (dccm_remap_create_jones99 / _bilinear).  Pairs with
    different longitudes come back in separable form (kind 2: longitude factors x latitude factors, multiplied out
    in the kernel); their results have the bits of the CSR operator built from the generated table, and nnz counts
    the entries that table holds.  Equal longitudes come back as zonal stencils (kind 1)."""
    import torch
    T = dccm.tables
    A, O, Sx = pair(orc, dccm, name)
    for label, s, d in (("O->S", O, Sx), ("S->O", Sx, O), ("A->O", A, O), ("O->A", O, A), ("A->S", A, Sx)):
        for cons in (True, False):
            tab = T.gen_table_jones99(s, d, 1, 1) if cons else T.gen_table_bilinear(s, d, 1)
            si, ri, cf = tab.index(s.im, d.im)
            csr = dccm.RemapOperator(si, ri, cf, s.n, d.n)                        # kind 0
            op = dccm.RemapOperator.from_grids(s, d, cons)
            assert csr.kind == 0 and op.kind == (1 if s.im == d.im else 2), (label, cons, op.kind)
            assert op.nnz == len(cf), (label, cons)
            for D in (1, 3, 10, 13):
                x = torch.as_tensor(S.generic_fields(np, s, D), device=gpu).contiguous()
                assert torch.equal(op.apply(x), csr.apply(x)), (label, cons, D)
            x = S.generic_fields(np, s, 4)
            assert np.array_equal(op.apply_host(x), orc.remap_apply(si, ri, cf, x, d.n)), (label, cons, "host")


@pytest.mark.parametrize("groups", [None, 3, 100])
@pytest.mark.parametrize("D", [1, 5, 8, 17, 43])
def test_remap_field_counts_and_zero_fill(gpu, orc, dccm, S, D, groups, monkeypatch):
    """recv_data(:,:) = 0 covers ALL rn2 columns and rows beyond the table
    (ref common/interpolation_data_latlon_mod.f90:293).  groups: the host form moves its fields in that many
    pipelined groups (H2D | kernel | D2H)."""
    if groups:
        monkeypatch.setenv("DCCM_HOST_CHUNKS", str(groups))
    A, O, Sx = pair(orc, dccm, "T21_Pl42")
    tab = dccm.tables.gen_table_jones99(A, Sx, 2)
    send_i, recv_i, coef = tab.index(A.im, Sx.im)
    op = dccm.RemapOperator(send_i, recv_i, coef, A.n, Sx.n)
    x = S.generic_fields(np, A, D + 2)
    got = op.apply_host(x, rn1=Sx.n + 7, rn2=D + 3, num_of_data=D)
    ref = orc.remap_apply(send_i, recv_i, coef, x, Sx.n + 7, D + 3, D)
    assert got.shape == (D + 3, Sx.n + 7)
    assert np.array_equal(got, ref)
    assert np.all(got[D:] == 0.0) and np.all(got[:, Sx.n:] == 0.0)


def test_remap_ragged_empty_and_duplicate_rows(gpu, orc, dccm):
    rng = np.random.default_rng(7)
    n_send, n_recv, nops = 300, 257, 2000
    send_i = rng.integers(1, n_send + 1, nops).astype(np.int32)
    recv_i = rng.integers(1, n_recv + 1, nops).astype(np.int32)
    recv_i[recv_i % 5 == 0] = 1                      # empty rows + one very long row
    coef = rng.standard_normal(nops)
    x = rng.standard_normal((3, n_send))
    op = dccm.RemapOperator(send_i, recv_i, coef, n_send, n_recv)
    assert np.array_equal(op.apply_host(x), orc.remap_apply(send_i, recv_i, coef, x, n_recv))
    # no operations at all: everything is zero-filled
    op0 = dccm.RemapOperator(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), 4, 6)
    assert np.all(op0.apply_host(np.ones((2, 4))) == 0.0)
    with pytest.raises(dccm.DccmError, match="out of range"):
        dccm.RemapOperator([1, 9], [1, 1], [1.0, 1.0], 4, 6)
    with pytest.raises(dccm.DccmError):
        op.apply_host(x, rn1=n_recv - 1)


def test_fortran_external_interpolate_data_symbol(gpu, orc, dccm, S):
    """`interpolate_data_` as Jcup calls it (ref common/interpolate_data.f90:1-17): every argument by reference,
    blank-padded component names with their hidden lengths; same bits as the oracle's loop."""
    import ctypes as C
    L = dccm._lib
    lib = dccm.lib()
    A, O, Sx = pair(orc, dccm, "T42_T42")
    send_i, recv_i, coef = dccm.tables.gen_table_jones99(A, Sx, 2).index(A.im, Sx.im)
    op = dccm.RemapOperator(send_i, recv_i, coef, A.n, Sx.n, A.im, Sx.im)
    L.check(lib.dccm_interp_set_model_name(1, b"ATM"))
    L.check(lib.dccm_interp_set_model_name(3, b"SFC"))
    L.check(lib.dccm_interp_register(3, 1, 2, op._h))
    x = S.generic_fields(np, A, 6)
    recv = np.full((7, Sx.n), np.nan)
    ints = [C.c_int32(v) for v in (2, A.n, 6, Sx.n, 7, 4, 1, 1)]          # mapping_tag sn1 sn2 rn1 rn2 num_of_data tn exchange_tag
    p = [C.cast(C.byref(v), L.i32p) for v in ints]
    L.f77_interpolate_data()(b"SFC       ", b"ATM   ", p[0], p[1], p[2], L.dp(x), p[3], p[4], L.dp(recv), p[5], p[6], p[7], 10, 6)
    assert np.array_equal(recv, orc.remap_apply(send_i, recv_i, coef, x, Sx.n, 7, 4))


@pytest.mark.parametrize("sizes", [(360, 181, 128, 65), (128, 65, 360, 181)])
def test_remap_with_the_standalone_regular_grid_table(gpu, orc, dccm, sizes):
    """Tables of common/cal_mappingtable.f90:10-49 (make_mapping_table) through the remap apply, bit-exact."""
    nx_r, ny_r, nx_s, ny_s = sizes
    send_i, recv_i, coef = dccm.tables.make_mapping_table(*sizes).index(nx_s, nx_r)
    x = np.random.default_rng(11).standard_normal((3, nx_s * ny_s))
    op = dccm.RemapOperator(send_i, recv_i, coef, nx_s * ny_s, nx_r * ny_r)
    assert np.array_equal(op.apply_host(x), orc.remap_apply(send_i, recv_i, coef, x, nx_r * ny_r))


def test_interpolate_data_callback_interface(gpu, orc, dccm, S):
    """The (recv_model, send_model, mapping_tag)-keyed path Jcup drives
    (ref common/interpolate_data.f90:1-17, interpolation_data_latlon_mod.f90:116-154,219-270)."""
    m = dccm.interpolation_data_latlon_mod
    A, O, Sx = pair(orc, dccm, "T42_T42")
    m.interpolation_data_latlon_Init(3, 2, 3)
    tabs = {1: dccm.tables.gen_table_bilinear(A, Sx), 2: dccm.tables.gen_table_jones99(A, Sx, 2)}
    for tag, tab in tabs.items():
        send_i, recv_i, coef = tab.index(A.im, Sx.im)
        m.set_operation_index("SFC", "ATM", tag, send_data_index=send_i, recv_data_index=recv_i,
                              num_of_send_grid=A.n, num_of_recv_grid=Sx.n)
        m.set_interpolate_coef("ATM", "SFC", "SFC", tag, coef)
    x = S.generic_fields(np, A, 8)
    for tag, tab in tabs.items():
        send_i, recv_i, coef = tab.index(A.im, Sx.im)
        recv = np.full((8, Sx.n), np.nan)
        dccm.interpolate_data("SFC", "ATM", tag, A.n, 8, x, Sx.n, 8, recv, 5, 1, np.array([1]))
        assert np.array_equal(recv, orc.remap_apply(send_i, recv_i, coef, x, Sx.n, 8, 5))
    with pytest.raises(dccm.DccmError, match="no operation index"):
        dccm.interpolate_data("ATM", "SFC", 1, Sx.n, 1, np.zeros((1, Sx.n)), A.n, 1, np.zeros((1, A.n)), 1)


def test_remap_device_resident_equals_host(gpu, orc, dccm, S):
    import torch
    A, O, Sx = pair(orc, dccm, "T106_1deg")
    tab = dccm.tables.gen_table_jones99(O, Sx, 1, 1)
    send_i, recv_i, coef = tab.index(O.im, Sx.im)
    op = dccm.RemapOperator(send_i, recv_i, coef, O.n, Sx.n)
    x = S.generic_fields(torch, O, 12, dev=gpu)
    y = op.apply(x)
    torch.cuda.synchronize()
    ref = orc.remap_apply(send_i, recv_i, coef, x.cpu().numpy(), Sx.n)
    assert np.array_equal(y.cpu().numpy(), ref)


# ------------------------------------------------------------------ K2 bulk flux

def test_branch_free_division_and_sqrt_match_the_operators_bitwise(gpu, dccm):
    """FastArith (the compiler's own fast-path sequences without the per-operation branch) against `/`, `1.0/x`
    and `sqrt` on the device: every accepted result has the operator's bits; operands with extreme exponents,
    zeros, infinities and NaN are rejected (and then take the IEEE re-evaluation), never wrong."""
    import ctypes as C
    import torch
    L = dccm._lib
    g = torch.Generator(device=gpu); g.manual_seed(1234)
    n = 1 << 22
    def draw(spread):
        m = torch.rand(n, generator=g, device=gpu, dtype=torch.float64) + 1.0
        e = torch.randint(-spread, spread + 1, (n,), generator=g, device=gpu)
        s = torch.randint(0, 2, (n,), generator=g, device=gpu, dtype=torch.float64) * 2.0 - 1.0
        return torch.ldexp(m * s, e)
    for spread, max_rej in ((40, 0), (300, None), (1070, None)):
        a, b = draw(spread), draw(spread)
        if spread == 1070:      # specials
            a[:8] = torch.tensor([0.0, -0.0, float("inf"), float("nan"), 1.0, 1.0, 5e-324, 1e308], device=gpu)
            b[:8] = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0, float("inf"), 3.0, 1e-308], device=gpu)
        bad, rej = C.c_int64(-1), C.c_int64(-1)
        L.check(L.lib().dccm_selftest_fast_arith_device(L.tptr(a), L.tptr(b), n, C.byref(bad), C.byref(rej)))
        print(f"exponent spread 2^+-{spread}: {rej.value} of {4 * n} operations rejected, {bad.value} mismatches")
        assert bad.value == 0
        if max_rej is not None:
            assert rej.value <= max_rej
        if spread == 1070:
            assert rej.value > 0


def _bulk_case(S, dccm, im, jm):
    from test_oracle_kat import _bulk_inputs
    g = dccm.tables.get_LonLatGrid(im, jm)
    return _bulk_inputs(S, g)


def _run_bulk_gpu(dccm, IA, JA, inp):
    d = dccm.dsfcm
    out = {k: np.full((3, JA, IA), np.nan) for k in d.OUT3}
    out["DelVarImplCPL"] = np.full((4, JA, IA), np.nan)
    out["SfcTemp"] = inp["SfcTemp"].copy()
    out["SfcAlbedo"] = inp["SfcAlbedo"].copy()
    dccm.DSFCM_Util_SfcBulkFlux_Get(
        IA, JA, out["WindStressX"], out["WindStressY"], out["SenHFlx"], out["QVapMFlx"], out["LatHFlx"],
        out["SfcVelTransCoef"], out["SfcTempTransCoef"], out["SfcQVapTransCoef"], out["DelVarImplCPL"],
        out["SUwRFlx"], out["LUwRFlx"], out["SfcHFlx_ns"], out["SfcHFlx_sr"], out["DSfcHFlxDTs"],
        inp["WindU"], inp["WindV"], inp["SfcAirTemp"], inp["QVap1"], inp["SDwRFlx"], inp["LDwRFlx"],
        inp["ImplCplCoef1"], inp["ImplCplCoef2"], out["SfcTemp"], out["SfcAlbedo"], inp["SIceCon"],
        inp["Sig1Info"], inp["SfcHeight"], inp["SfcPress"])
    return out


def _check_bulk(got, ref):
    """every output of DSFCM_Util_SfcBulkFlux_Get, halo included (NaN = never written): identical bits."""
    from exchange_ref import bits_equal
    for k in ref:
        assert bits_equal(got[k], ref[k]), f"{k}: rel err {relerr(got[k], ref[k])}"


def test_portable_exp_log_pow_device_equals_oracle_bitwise(gpu, orc, dccm):
    """csrc/dccm_pmath.cuh on the device against oracle/orc_pmath.h on the host, 4 M operands per function over
    the whole exponent range plus the special values; the division inside log goes through FastArith."""
    import torch
    L = dccm._lib
    rng = np.random.default_rng(77)
    n = 1 << 22
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, 1e-310, 2.2250738585072014e-308, 1.0, -1.0,
                        709.78, 709.79, -745.0, -745.2, -708.4, -740.0, 1.7976931348623157e308])
    kappa = 8.3144621 / 1.8e-2 / 1616.0
    cases = [(0, np.concatenate([rng.uniform(-750, 720, n), rng.uniform(-1, 1, n), special]), 0.0, orc.pm_exp),
             (1, np.concatenate([np.exp(rng.uniform(-745, 709, n)), rng.uniform(0.5, 2.0, n), special]), 0.0, orc.pm_log),
             (2, np.concatenate([rng.uniform(0.3, 3.0, n), np.exp(rng.uniform(-30, 30, n)), special]), kappa,
              lambda x: orc.pm_pow(x, kappa)),
             (3, np.concatenate([rng.uniform(1e8, 1e11, n), np.exp(rng.uniform(-700, 700, n)), special]), 0.0,
              lambda x: orc.pm_pow(x, 0.25))]
    for which, x, y, ref in cases:
        xd = torch.as_tensor(x, device=gpu)
        out = torch.empty_like(xd)
        L.check(L.lib().dccm_selftest_pmath_device(which, L.tptr(xd), xd.numel(), y, L.tptr(out)))
        got, want = out.cpu().numpy(), ref(x)
        bad = got.view(np.int64) != want.view(np.int64)
        assert not bad.any(), (which, x[bad][:5], got[bad][:5], want[bad][:5])


@pytest.mark.parametrize("chunks", [None, 5, 1000])
def test_bulkflux_vs_oracle_T42(gpu, orc, dccm, S, chunks, monkeypatch):
    """chunks: the host entry point moves the interior rows in that many pipelined pieces (H2D | kernel | D2H)."""
    if chunks:
        monkeypatch.setenv("DCCM_HOST_CHUNKS", str(chunks))
    IA, JA, inp = _bulk_case(S, dccm, 128, 64)
    got = _run_bulk_gpu(dccm, IA, JA, inp)
    ref = orc.bulkflux(IA, JA, inp)
    _check_bulk(got, ref)
    # halo cells are never written, slot 3 of SfcTemp/SfcAlbedo halo keeps the caller's value
    assert np.all(np.isnan(got["SenHFlx"][:, 0, :])) and np.all(np.isnan(got["LUwRFlx"][:, :, 0]))
    assert np.all(got["SfcTemp"][2][0, :] == -999.0)
    # ice-free columns: the sea-ice slot is exactly zero (CalcFlag false, ref :268-272, :339-347)
    ice0 = inp["SIceCon"][1:-1, 1:-1] == 0.0
    assert ice0.any() and np.all(got["SenHFlx"][1][1:-1, 1:-1][ice0] == 0.0)


def test_bulkflux_vs_golden(gpu, orc, dccm, S):
    IA, JA, inp = _bulk_case(S, dccm, 16, 8)
    got = _run_bulk_gpu(dccm, IA, JA, inp)
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "bulkflux_16x8.npz")) as z:
        _check_bulk({k: np.ascontiguousarray(got[k][:, 1:-1, 1:-1]) for k in z.files}, {k: z[k] for k in z.files})


def test_bulkflux_extreme_columns(gpu, orc, dccm, S):
    """calm wind (velocity floors, ref :576-582), strongly stable / unstable stratification
    (both Louis branches, :490-534, coefficient clamps :555-565), ice fraction at the 1e-12 test."""
    IA, JA, inp = _bulk_case(S, dccm, 16, 8)
    I = (slice(1, -1), slice(1, -1))
    inp["WindU"][I][0, :] = 0.0; inp["WindV"][I][0, :] = 0.0
    inp["WindU"][I][1, :] = 1e-4; inp["WindV"][I][1, :] = -1e-4
    inp["SfcAirTemp"][I][2, :] = inp["SfcTemp"][0][I][2, :] + 25.0     # very stable
    inp["SfcAirTemp"][I][3, :] = inp["SfcTemp"][0][I][3, :] - 25.0     # very unstable
    inp["WindU"][I][4, :] = 900.0; inp["WindV"][I][4, :] = 900.0       # above VelMax
    inp["SIceCon"][I][5, :] = 1e-12                                    # not > 1e-12: no ice
    inp["SIceCon"][I][6, :] = 1.0000001e-12                            # just above
    inp["SIceCon"][I][7, :] = 1.0
    got = _run_bulk_gpu(dccm, IA, JA, inp)
    ref = orc.bulkflux(IA, JA, inp)
    _check_bulk(got, ref)
    assert np.all(got["SfcHFlx_ns"][1][I][5, :] == 0.0) and np.all(got["SfcHFlx_ns"][1][I][6, :] != 0.0)


# ------------------------------------------------------------------ K3/K4 column solves

def _vdiff_case(S, dccm, im, jm, K, nc):
    g = dccm.tables.get_LonLatGrid(im, jm)
    return g, S.column_inputs(np, g, K, nc)


@pytest.mark.parametrize("im,jm,K,nc,iq", [(128, 64, 26, 1, 1), (64, 32, 16, 3, 2), (8, 4, 2, 1, 1), (9, 5, 3, 4, 4), (16, 8, 9, 7, 5)])
def test_vdiff_reference_order_mode_is_bit_exact(gpu, orc, dccm, S, im, jm, K, nc, iq):
    g, inp = _vdiff_case(S, dccm, im, jm, K, nc)
    args = (g.im, g.jm, K, nc, iq, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    ref_h = orc.VDiff(*args)
    ref = ref_h.forward(inp)
    h = dccm.SfcImplicitCoupling(*args)
    got = h.VDiffForward(inp)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), f"forward {k}: rel err {relerr(got[k], ref[k])}"
    # the surface component hands back the level-1 increment (ref atm/dccm_atm_mod.f90:832-835)
    lvl1 = 1e-3 * np.stack([S.normal(np, np.arange(g.n, dtype=np.float64), 70.0 + k) for k in range(4)])
    DU, DV, DT, DQ = (got[k].copy() for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"))
    DU[0], DV[0], DT[0], DQ[iq - 1, 0] = lvl1
    refb = ref_h.backward(DU, DV, DT, DQ)
    h.VDiffBackward(DU, DV, DT, DQ)
    for a, b, k in zip((DU, DV, DT, DQ), refb, ("DUDt", "DVDt", "DTempDt", "DQMixDt")):
        assert np.array_equal(a, b), f"backward {k}: rel err {relerr(a, b)}"


def test_vdiff_redo_list_takes_the_columns_the_branch_free_division_rejects(gpu, orc, dccm, S):
    """Reference-order forward solve on columns whose operands leave the range of the branch-free division sequence
    (fluxes of 1e-300 -> numerators below 2^-969; a zero diffusion coefficient -> division by zero): those columns go
    to the redo list and are solved again with the plain IEEE operators, the others keep the fast path; every finite
    value has the oracle's bits, Inf / NaN appear in the same places (NaN payloads differ between x86 and the GPU)."""
    g, inp = _vdiff_case(S, dccm, 64, 32, 12, 2)
    inp = {k: np.array(v, copy=True) for k, v in inp.items()}
    tiny, zero = slice(100, 164), slice(700, 764)
    inp["MomFluxX"][:, tiny] *= 1e-300
    inp["HeatFlux"][:, tiny] *= 1e-290
    inp["TempDiffCoef"][:, zero] = 0.0
    args = (g.im, g.jm, 12, 2, 1, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    with np.errstate(all="ignore"):
        ref = orc.VDiff(*args).forward(inp)
    h = dccm.SfcImplicitCoupling(*args)
    assert h.redo_total() == 0
    got = h.VDiffForward(inp)
    n = h.redo_total()
    print("columns redone with IEEE operators:", n, "of", g.n)
    assert 128 <= n < g.n // 2
    for k in ref:
        a, b = got[k], ref[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        m = ~np.isnan(b)
        assert np.array_equal(a[m].view(np.int64), b[m].view(np.int64)), f"{k}: rel err {relerr(a[m], b[m])}"
    assert np.isfinite(ref["DUDt"][:, tiny]).all() and (ref["DUDt"][1:, tiny] != 0).any()
    # a second, ordinary call on the same handle: the list was cleared, nothing is redone
    g2, ok = _vdiff_case(S, dccm, 64, 32, 12, 2)
    got2 = h.VDiffForward(ok)
    assert h.redo_total() == n
    ref2 = orc.VDiff(*args).forward(ok)
    for k in ref2:
        assert np.array_equal(got2[k], ref2[k]), k


@pytest.mark.parametrize("chunks", [3, 7, 1000])
def test_vdiff_host_calls_pipelined_over_column_chunks(gpu, orc, dccm, S, chunks, monkeypatch):
    """The *_host entry points move their arguments in column chunks (H2D | kernel | D2H on three streams, 2-D
    copies of `levels` rows); forcing 3, 7 (ragged) and more chunks than 32-column groups on a small grid gives
    the bits of the oracle, for pinned and for pageable host arrays alike."""
    import torch
    g, inp = _vdiff_case(S, dccm, 72, 37, 9, 3)
    args = (g.im, g.jm, 9, 3, 2, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    ref_h = orc.VDiff(*args)
    ref = ref_h.forward(inp)
    monkeypatch.setenv("DCCM_HOST_CHUNKS", str(chunks))
    h = dccm.SfcImplicitCoupling(*args)
    pinned = {k: torch.as_tensor(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in inp.items()}
    for host_in in (inp, pinned):
        got = h.VDiffForward(host_in)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), k
    DU, DV, DT, DQ = (got[k].copy() for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"))
    DU[0] = 1e-3; DV[0] = -2e-3; DT[0] = 5e-4; DQ[1, 0] = 1e-7
    refb = ref_h.backward(DU.copy(), DV.copy(), DT.copy(), DQ.copy())
    h.VDiffBackward(DU, DV, DT, DQ)
    for a, b in zip((DU, DV, DT, DQ), refb):
        assert np.array_equal(a, b)


def test_host_register_pins_caller_arrays(gpu, dccm):
    """dccm_host_register / _unregister: page-lock a caller-owned array (what a Fortran component does once at init
    for its module arrays) -- the array is then usable by the pipelined *_host entry points like any other."""
    import torch
    L = dccm._lib
    a = np.arange(1 << 16, dtype=np.float64)
    L.check(L.lib().dccm_host_register(a.ctypes.data, a.nbytes))
    try:
        t = torch.empty(a.size, dtype=torch.float64, device=gpu)
        t.copy_(torch.from_numpy(a), non_blocking=True)
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), a)
    finally:
        L.check(L.lib().dccm_host_unregister(a.ctypes.data))
    with pytest.raises(dccm.DccmError):
        L.check(L.lib().dccm_host_register(None, 8))


def test_vdiff_fast_mode_within_tolerance(gpu, orc, dccm, S):
    g, inp = _vdiff_case(S, dccm, 128, 64, 26, 2)
    args = (g.im, g.jm, 26, 2, 1, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    ref_h = orc.VDiff(*args)
    ref = ref_h.forward(inp)
    h = dccm.SfcImplicitCoupling(*args, fast=True)
    got = h.VDiffForward(inp)
    for k in ref:
        assert relerr(got[k], ref[k]) <= RTOL, k
    DU, DV, DT, DQ = (got[k].copy() for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"))
    rb = ref_h.backward(ref["DUDt"], ref["DVDt"], ref["DTempDt"], ref["DQMixDt"])
    h.VDiffBackward(DU, DV, DT, DQ)
    for a, b in zip((DU, DV, DT, DQ), rb):
        assert relerr(a, b, floor=1e-9 * np.abs(b).max()) <= RTOL


def test_vdiff_vs_golden(gpu, orc, dccm, S):
    g, inp = _vdiff_case(S, dccm, 8, 4, 6, 2)
    h = dccm.SfcImplicitCoupling(g.im, g.jm, 6, 2, 2, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    got = h.VDiffForward(inp)
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "vdiff_8x4_K6.npz")) as z:
        for k in got:
            assert relerr(got[k], z["fwd_" + k]) <= 1e-13, k


def test_vdiff_argument_errors(gpu, dccm):
    with pytest.raises(dccm.DccmError, match="kmax"):
        dccm.SfcImplicitCoupling(4, 2, 1, 1, 1, 9.8, 1004.6, 287.04, 1200.0)
    with pytest.raises(dccm.DccmError, match="IndexH2OVap"):
        dccm.SfcImplicitCoupling(4, 2, 5, 2, 3, 9.8, 1004.6, 287.04, 1200.0)


# ------------------------------------------------------------------ full-size properties

def test_full_size_properties_T1279(gpu, dccm, S):
    """BASELINE config 5 sizes, checked through size-independent properties on the device:
    a constant field is preserved by the conservative and bilinear A->S tables, remap is linear,
    and the column solve satisfies its own tridiagonal system (residual)."""
    import torch
    T = dccm.tables
    A = T.get_LonLatGrid(3840, 1920)
    O = T.regular_LonLatGrid(3600, 1800)
    Sx = T.generate_surface_exchange_grid(A, O)
    assert (Sx.im, Sx.jm) == (3840, 3718)
    for tab in (T.gen_table_jones99(A, Sx, 1, 1), T.gen_table_bilinear(A, Sx, 1)):
        send_i, recv_i, coef = tab.index(A.im, Sx.im)
        op = dccm.RemapOperator(send_i, recv_i, coef, A.n, Sx.n)
        del send_i, recv_i, coef
        x = torch.full((2, A.n), 3.5, dtype=torch.float64, device=gpu)
        x[1] = S.generic_fields(torch, A, 1, dev=gpu)[0]
        y = op.apply(x)
        assert float((y[0] / 3.5 - 1.0).abs().max()) <= 1e-12
        y2 = op.apply(2.0 * x)                                  # exact in binary fp
        assert torch.equal(y2, 2.0 * y)
        del op, x, y, y2
    # column solve on a 1/8 latitude band of the T1279 atmosphere
    K, nc, j1 = 26, 1, 240
    inp = S.column_inputs(torch, A, K, nc, 0, j1, dev=gpu)
    ncol = j1 * A.im
    h = dccm.SfcImplicitCoupling(A.im, j1, K, nc, 1, S.GRAV, S.CPDRY, S.GASRDRY, S.DELTIME)
    out = {"DUDt": torch.empty((K, ncol), dtype=torch.float64, device=gpu),
           "DVDt": torch.empty((K, ncol), dtype=torch.float64, device=gpu),
           "DTempDt": torch.empty((K, ncol), dtype=torch.float64, device=gpu),
           "DQMixDt": torch.empty((nc, K, ncol), dtype=torch.float64, device=gpu),
           "ImplCplCoef1": torch.empty((4, ncol), dtype=torch.float64, device=gpu),
           "ImplCplCoef2": torch.empty((4, ncol), dtype=torch.float64, device=gpu)}
    h.forward_device(inp, out)
    # choose x1 from the reduced equation with a synthetic surface coefficient, then back-substitute
    C = 0.012
    x1 = out["ImplCplCoef2"][0] / (out["ImplCplCoef1"][0] + C)
    lvl1 = torch.stack([x1, x1, x1, x1])
    h.backward_device(out, lvl1)
    torch.cuda.synchronize()
    x = out["DUDt"] * (2.0 * S.DELTIME)
    P, Tv, H, D, FX = inp["Press"], inp["VirTemp"], inp["Height"], inp["VelDiffCoef"], inp["MomFluxX"]
    Tc = torch.zeros((K + 1, ncol), dtype=torch.float64, device=gpu)
    Tc[1:K] = D[1:K] * (P[1:K] / (S.GASRDRY * Tv[1:K]) / (H[1:K] - H[0:K - 1]))
    m = -(P[1:] - P[:-1]) / S.GRAV / (2.0 * S.DELTIME)
    r = -(FX[1:] - FX[:-1])
    res = (m + Tc[:-1] + Tc[1:]) * x - r
    res[1:] -= Tc[1:K] * x[:-1]
    res[:-1] -= Tc[1:K] * x[1:]
    res[0] += C * x[0]
    scale = (m * x).abs().max()
    assert float(res.abs().max() / scale) <= 1e-11


# ------------------------------------------------------------------ whole exchange step

@pytest.mark.parametrize("name,K,fast,order", [("T21_Pl42", 16, False, 1), ("T42_T42", 26, False, 1),
                                               ("T21_1deg", 26, False, 1), ("T106_1deg", 26, False, 1),
                                               ("T42_T42", 26, False, 2), ("T21_1deg", 26, True, 1)])
def test_exchange_step_vs_oracle(gpu, orc, dccm, S, name, K, fast, order):
    """Whole exchange against the oracle.  Reference-order mode: EVERY stage output -- coupling coefficients, the
    22 remapped surface inputs, the 21 put-side layers (bulk flux included), the remapped fluxes on the atmosphere
    and ocean grids and the four tendencies -- has the oracle's bits.  `fast` (shared reciprocals in the forward
    solve, not the default) is held to 1e-12 on the coefficients and to the conditioning-aware bar downstream."""
    import torch
    from exchange_ref import compare_exchange, floor_rel, oracle_exchange
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, name)
    tabs = X.build_tables(A, O, Sx, order_as=order)
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, fast=fast, device=gpu)
    col, atm, ocn = S.column_inputs(np, A, K, 1), S.atm_surface_fields(np, A), S.ocn_surface_fields(np, O)
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    ex.set_inputs(tt(col), {k: v[None] for k, v in tt(atm).items()}, {k: v[None] for k, v in tt(ocn).items()})
    ex.step(fused=False)
    torch.cuda.synchronize()
    ref = oracle_exchange(orc, S, A, O, Sx, K, 1, 1, tabs, col, atm, ocn)
    detail, same = {}, {}
    worst = compare_exchange(ex, ref, detail=detail, bitwise=same)
    print(name, "fast" if fast else "reference-order", {k: float("%.2e" % v) for k, v in detail.items()}, same)
    if not fast:
        assert all(same.values()), {k: detail[k] for k, v in same.items() if not v}
        assert worst == 0.0
    else:
        tol = {k: RTOL for k in detail}
        for fld in ("SfcPress", "SfcAirTemp"):          # the oracle's own response to 1-ulp input changes
            atm2 = dict(atm)
            atm2[fld] = atm[fld] * (1.0 + 2.0 ** -52)
            ref2 = oracle_exchange(orc, S, A, O, Sx, K, 1, 1, tabs, col, atm2, ocn)
            for k in ("s2a", "s2o", "a_recv", "o_recv"):
                tol[k] = max(tol[k], 4.0 * floor_rel(ref2[k], ref[k]))
            for k in ("DUDt", "DVDt", "DTempDt", "DQMixDt"):
                tol["bwd_" + k] = max(tol["bwd_" + k], 4.0 * floor_rel(ref2["bwd"][k], ref["bwd"][k]))
        for k in ("s_cons", "s_obil", "s_ocons"):
            assert same[k], k
        over = {k: (detail[k], tol[k]) for k in detail if detail[k] > tol[k]}
        assert not over, over
        assert worst <= 1e-11
    # the fused surface kernel (remap + bulk flux + pack in registers) gives the same bits
    keep = {k: getattr(ex, k).clone() for k in ("s2a", "s2o", "a_recv", "o_recv")}
    keep.update({k: v.clone() for k, v in ex.tend.items()})
    ex.s2a.zero_(); ex.s2o.zero_()
    ex.step(fused=True)
    torch.cuda.synchronize()
    for k in ("s2a", "s2o", "a_recv", "o_recv"):
        assert torch.equal(getattr(ex, k), keep[k]), k
    for k in ex.tend:
        assert torch.equal(ex.tend[k], keep[k]), k
    # global integrals of the conservatively remapped fluxes agree to the conservation bar
    wA = np.repeat(A.y_LatWt, A.im)
    got, want = ex.a_recv[:4].cpu().numpy(), ref["a_recv"][:4]
    assert np.abs((got * wA).sum(1) / (want * wA).sum(1) - 1.0).max() <= CONS


@pytest.mark.parametrize("name,members", [("T21_Pl42", 1), ("T42_T42", 3), ("T21_1deg", 1), ("T106_1deg", 2)])
def test_fused_surface_kernel_staged_and_direct_forms_bit_exact(gpu, orc, dccm, S, name, members):
    """The TMA-staged form of the fused surface kernel (atmosphere rows in shared memory) and the direct form
    (per-thread global gathers) against the unfused remap -> bulk flux -> pack sequence: same bits in the 21
    put-side layers and in the API-complete DSFCM arrays; rows shorter and longer than a CTA, wrap-around
    pieces (T21: one CTA holds a whole latitude circle), ragged last CTA (T106: 320 = 2.5 CTAs), ensembles."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    L = dccm._lib
    A, O, Sx = pair(orc, dccm, name)
    tabs = X.build_tables(A, O, Sx)
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
    want_full = {k: v.clone() for k, v in ex.sfc_out.items()}
    want_full["s_obil"] = ex.s_obil.clone(); want_full["s_ocons"] = ex.s_ocons.clone()
    forms = []
    try:
        for staged in (1, 0):
            for minb in (5, 4, 6):
                for full in (False, True):
                    ex.configure_sfc(staged, minb)
                    ex.s2a.fill_(float("nan")); ex.s2o.fill_(float("nan"))
                    if full:
                        for v in ex.sfc_out.values():
                            v.fill_(float("nan"))
                        ex.s_obil[2 * M:].fill_(float("nan")); ex.s_ocons[3 * M:].fill_(float("nan"))
                    ex.sfc_fused(store_full=full)
                    torch.cuda.synchronize()
                    forms.append(ex.sfc_last_form())
                    tag = f"staged={staged} minb={minb} full={full}"
                    assert torch.equal(ex.s2a, want["s2a"]), tag + " s2a"
                    assert torch.equal(ex.s2o, want["s2o"]), tag + " s2o"
                    if full:
                        for k, v in ex.sfc_out.items():
                            rows = 2 * M if k in ("SfcHFlx_ns", "SfcHFlx_sr", "DSfcHFlxDTs") else v.shape[0]   # 2-slot arrays
                            assert torch.equal(v[:rows], want_full[k][:rows]), tag + " " + k
                        assert torch.equal(ex.s_obil, want_full["s_obil"]), tag
                        assert torch.equal(ex.s_ocons, want_full["s_ocons"]), tag
    finally:
        ex.configure_sfc(1, 5)
    # every A->S table of these grid pairs is a zonal stencil on even longitudes: the staged form must have run
    assert forms[:6] == [1] * 6 and forms[6:] == [0] * 6, forms


@pytest.mark.parametrize("how", ["few", "all"])
def test_fused_surface_kernel_redo_list_for_out_of_range_operands(gpu, orc, dccm, S, how):
    """Cells whose divisions leave the exponent range of the branch-free arithmetic (denormal pressure here) are
    listed by the fused kernel and re-evaluated with the plain IEEE operators by the redo kernel: same bits --
    infinities and NaN included -- as the unfused sequence.  "all": more cells than the list holds (every cell
    is redone); a second, clean call afterwards shows the list was cleared."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    L = dccm._lib
    A, O, Sx = pair(orc, dccm, "T106_1deg")
    tabs = X.build_tables(A, O, Sx)
    K = 8
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    col, atm, ocn = tt(S.column_inputs(np, A, K, 1)), tt(S.atm_surface_fields(np, A)), tt(S.ocn_surface_fields(np, O))
    clean = {k: v.clone() for k, v in atm.items()}
    if how == "few":
        atm["SfcPress"][1000:1003] = 1e-310
        atm["WindU"][5000] = 0.0; atm["WindV"][5000] = 0.0        # calm: sqrt(0) stays on the fast path
        atm["SfcAirTemp"][20000] = float("inf")
    else:
        atm["SfcPress"][:] = 1e-310
    same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=1.25e300), torch.nan_to_num(b, nan=1.25e300))
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, device=gpu)
    try:
        for inputs in (atm, clean):
            ex.set_inputs(col, {k: v[None] for k, v in inputs.items()}, {k: v[None] for k, v in ocn.items()})
            ex.forward()
            ex.remap_to_sfc(); ex.bulk(); ex.pack_sfc()
            torch.cuda.synchronize()
            want = (ex.s2a.clone(), ex.s2o.clone())
            assert inputs is clean or not bool(torch.isfinite(want[1]).all())
            for staged in (1, 0):
                ex.configure_sfc(staged, 5)
                ex.s2a.zero_(); ex.s2o.zero_()
                ex.sfc_fused()
                torch.cuda.synchronize()
                assert same(ex.s2a, want[0]) and same(ex.s2o, want[1]), (how, staged, inputs is clean)
    finally:
        ex.configure_sfc(1, 5)


@pytest.mark.parametrize("name,members", [("T21_1deg", 1), ("T106_1deg", 2), ("T42_T42", 1)])
def test_exchange_from_grids_equals_exchange_from_tables(gpu, orc, dccm, S, name, members):
    """SurfaceExchange built straight from the grids (no table on the host; ocean-side operators separable where the
    longitudes differ) against the same exchange built from the generated tables: every output bit for bit, fused
    (staged and direct form) and unfused."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    L = dccm._lib
    A, O, Sx = pair(orc, dccm, name)
    M, K = members, 9
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    cols = [tt(S.column_inputs(np, A, K, 1, member=m)) for m in range(M)]
    atms = [tt(S.atm_surface_fields(np, A, member=m)) for m in range(M)]
    ocns = [tt(S.ocn_surface_fields(np, O, member=m)) for m in range(M)]
    inputs = ({k: torch.cat([c[k] for c in cols], dim=-1).contiguous() for k in cols[0]},
              {k: torch.stack([a[k] for a in atms]) for k in atms[0]},
              {k: torch.stack([o[k] for o in ocns]) for k in ocns[0]})
    ref = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=X.build_tables(A, O, Sx), members=M, device=gpu)
    ref.set_inputs(*inputs)
    ref.step(fused=False)
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, members=M, device=gpu)
    want_kind = 2 if O.im != Sx.im else 1
    assert {ex.ops[k].kind for k in ("os_cons", "os_bil", "so_cons", "so_bil")} == {want_kind}
    assert ex.nnz == ref.nnz
    ex.set_inputs(*inputs)
    try:
        for fused, staged in ((False, 1), (True, 1), (True, 0)):
            ex.configure_sfc(staged, 5)
            for t in (ex.s2a, ex.s2o, ex.a_recv, ex.o_recv):
                t.fill_(float("nan"))
            ex.step(fused=fused)
            torch.cuda.synchronize()
            for k in ("s2a", "s2o", "a_recv", "o_recv"):
                assert torch.equal(getattr(ex, k), getattr(ref, k)), (k, fused, staged)
            for k in ex.tend:
                assert torch.equal(ex.tend[k], ref.tend[k]), (k, fused, staged)
    finally:
        ex.configure_sfc(1, 5)


def test_exchange_ensemble_members_match_single_runs(gpu, orc, dccm, S):
    """BASELINE config 2: members batched along the layer axis share the tables; member m of the
    batch equals a single-member run on member m's inputs, bit for bit."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, Sx = pair(orc, dccm, "T21_Pl42")
    tabs = X.build_tables(A, O, Sx)
    K, M = 16, 3
    tt = lambda d: {k: torch.as_tensor(v, device=gpu).contiguous() for k, v in d.items()}
    cols = [tt(S.column_inputs(np, A, K, 1, member=m)) for m in range(M)]
    atms = [tt(S.atm_surface_fields(np, A, member=m)) for m in range(M)]
    ocns = [tt(S.ocn_surface_fields(np, O, member=m)) for m in range(M)]
    exM = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, members=M, device=gpu)
    exM.set_inputs({k: torch.cat([c[k] for c in cols], dim=-1).contiguous() for k in cols[0]},
                   {k: torch.stack([a[k] for a in atms]) for k in atms[0]},
                   {k: torch.stack([o[k] for o in ocns]) for k in ocns[0]})
    exM.step()
    unf = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, members=M, device=gpu)
    unf.col_in, unf.a2s_bil, unf.a2s_cons, unf.o2s_bil, unf.o2s_cons = exM.col_in, exM.a2s_bil.clone(), exM.a2s_cons, exM.o2s_bil, exM.o2s_cons
    unf.step(fused=False)
    assert torch.equal(unf.o_recv, exM.o_recv) and torch.equal(unf.a_recv, exM.a_recv)
    for m in range(M):
        ex1 = X.SurfaceExchange(A, O, Sx, K, 1, 1, tabs=tabs, members=1, device=gpu)
        ex1.set_inputs(cols[m], {k: v[None] for k, v in atms[m].items()}, {k: v[None] for k, v in ocns[m].items()})
        ex1.step()
        torch.cuda.synchronize()
        assert torch.equal(exM.o_recv[m::M], ex1.o_recv)
        assert torch.equal(exM.a_recv[m::M], ex1.a_recv)
        assert torch.equal(exM.tend["DTempDt"][:, m * A.n:(m + 1) * A.n], ex1.tend["DTempDt"])


def test_sharded_exchange_two_gpus_bit_exact(gpu):
    """2 ranks over NCCL: each band of the sharded run equals the unsharded run (tests/sharded_gpu_check.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(here, "sharded_gpu_check.py"), "T106_1deg"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_host_pipelined_exchange_equals_resident(gpu, orc, dccm, S):
    """exchange_host.HostPipelinedExchange (host inputs -> host outputs, latitude slabs pipelined over
    three streams, halo rows copied between slabs) gives the bits of the resident single-shot step."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    XH = importlib.import_module("dennou-ccm_b200.exchange_host")
    A, O, Sx = pair(orc, dccm, "T106_1deg")
    K, nc, nslab = 12, 2, 5
    ex = X.SurfaceExchange(A, O, Sx, K, nc, 1, device=gpu)
    ex.set_inputs(S.column_inputs(torch, A, K, nc, dev=gpu),
                  {k: v[None] for k, v in S.atm_surface_fields(torch, A, dev=gpu).items()},
                  {k: v[None] for k, v in S.ocn_surface_fields(torch, O, dev=gpu).items()})
    ex.step()
    hx = XH.HostPipelinedExchange(A, O, Sx, K, nc, 1, nslab=nslab, device=gpu)
    for s in range(nslab):
        (a0, a1), (o0, o1) = hx.bands(s)
        # same generator, same device as the resident run: the fields are pure functions of the global
        # cell index, so the slab inputs are the very same bits
        for k, t in S.column_inputs(torch, A, K, nc, a0, a1, dev=gpu).items():
            hx.h_in[s][k].copy_(t)
        for k, t in S.atm_surface_fields(torch, A, a0, a1, dev=gpu).items():
            hx.h_in[s]["a:" + k].copy_(t)
        for k, t in S.ocn_surface_fields(torch, O, o0, o1, dev=gpu).items():
            hx.h_in[s]["o:" + k].copy_(t)
    for _ in range(2):                       # twice: stream hand-over between consecutive exchanges
        hx.step()
    hx.synchronize()
    cat = lambda k, dim: torch.cat([hx.h_out[s][k] for s in range(nslab)], dim=dim)
    for k, dim, ref in (("o_recv", 1, ex.o_recv), ("a_recv", 1, ex.a_recv), ("DUDt", 1, ex.tend["DUDt"]),
                        ("DQMixDt", 2, ex.tend["DQMixDt"])):
        assert torch.equal(cat(k, dim), ref.cpu()), k


def test_ocn_glue_kernels_bit_exact(gpu, orc, dccm, S):
    """the ocean / sea-ice glue's element-wise work on the device, writing the O->S send
    layers in place and assembling the S->O results (ref ocn/dccm_ocn_mod.f90:825-851, :978-993)."""
    import torch
    O = dccm.tables.regular_LonLatGrid(360, 180)
    f = S.ocn_surface_fields(np, O)
    n, ld = O.n, O.n + 64
    tsC = f["SfcTempI"] - 273.15
    IceMaskMin = 0.05
    want_ti, want_ai = orc.ocn_put_assemble(f["SfcTempO"], f["SfcAlbedoO"], f["SIceCon"], tsC, f["SfcAlbedoI"], IceMaskMin, 273.15)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=gpu)
    bil = torch.full((2, ld), float("nan"), dtype=torch.float64, device=gpu)
    cons = torch.full((3, ld), float("nan"), dtype=torch.float64, device=gpu)
    dccm.dccm_ocn_mod.ocn_put_assemble(t(f["SfcTempO"]), t(f["SfcAlbedoO"]), t(f["SIceCon"]), t(tsC), t(f["SfcAlbedoI"]),
                                       IceMaskMin, 273.15, bil, cons)
    assert np.array_equal(bil[0, :n].cpu().numpy(), f["SfcTempO"]) and np.array_equal(bil[1, :n].cpu().numpy(), want_ti)
    assert np.array_equal(cons[0, :n].cpu().numpy(), f["SIceCon"]) and np.array_equal(cons[2, :n].cpu().numpy(), want_ai)
    assert torch.isnan(bil[:, n:]).all()
    assert ((f["SIceCon"] >= IceMaskMin) != (f["SIceCon"] > 0)).any()       # the mask threshold matters here
    r = S.generic_fields(np, O, 12)
    got = dccm.dccm_ocn_mod.ocn_get_assemble(t(r), 1000.0)
    want = orc.ocn_get_assemble(r[0], r[1], r[10], r[2], r[3], r[4], r[5], r[6], 1000.0)
    for k in want:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


def _atm_sfcflx_inputs(orc, n, seed=7):
    rng = np.random.default_rng(seed)
    f = {k: rng.normal(0.0, 1.0, n) for k in orc.ATM_SFCFLX_IN}
    f["ExnerR0"] = 1.0 + 0.01 * rng.random(n); f["ExnerZ1"] = 0.99 + 0.01 * rng.random(n)
    f["TempN1"] = 280.0 + 10.0 * rng.normal(size=n); f["SnowFrac"] = np.clip(rng.normal(0.3, 0.5, n), 0.0, 1.0)
    for k in ("SurfVelTransCoef", "SurfTempTransCoef", "SurfQVapTransCoef"):
        f[k] = 0.01 + 0.02 * rng.random(n)
    f["SurfHumidCoef"] = np.ones(n)
    return f


def test_atm_surface_flux_bookkeeping_bit_exact(gpu, orc, dccm):
    """dcpam_StoreAtmSurfFlxInfo (ref atm/dcpam_main_mod.f90:1068-1112) on the device."""
    import torch
    n = 128 * 64 + 37
    f = _atm_sfcflx_inputs(orc, n)
    want = orc.atm_store_surf_flx(f, 2.5e6, 1004.6, 1200.0)
    got = dccm.dcpam_main_mod.dcpam_StoreAtmSurfFlxInfo({k: torch.as_tensor(v, device=gpu) for k, v in f.items()},
                                                         2.5e6, 1004.6, 1200.0)
    for k in want:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    with pytest.raises(dccm.DccmError):
        L = dccm._lib
        L.check(L.lib().dccm_atm_store_surf_flx_device(n, L.AtmSfcFlx(), 1.0, 1.0, 1.0, None))     # NULL fields


def test_time_average_of_surface_puts_bit_exact(gpu, orc, dccm, S):
    """Jcup RECV_MODE='AVG' of the S->O layers: device accumulate / finish against the oracle, two intervals."""
    import torch
    O = dccm.tables.regular_LonLatGrid(72, 36)
    avg = None
    for interval, nput in enumerate((3, 1)):
        puts = [S.generic_fields(np, O, 12) * (1.0 + 0.1 * k + interval) for k in range(nput)]
        want = orc.time_average(puts)
        for x in puts:
            t = torch.as_tensor(np.ascontiguousarray(x), device=gpu)
            avg = avg or dccm.dccm_ocn_mod.TimeAverage(t)
            avg.put(t)
        assert np.array_equal(avg.get().cpu().numpy(), want)


@pytest.mark.parametrize("name", ["T21_Pl42", "T42_T42"])
def test_legacy_two_component_exchange_bit_exact(gpu, orc, dccm, S, name):
    """the legacy A<->O topology the shipped exp/ configs run (12 a2o + 4 o2a layers,
    ref common/mod_common_params.f90:76-175) goes through the same K1; T21 <-> axisymmetric Pl42 is the
    shipped grid pair (ref exp/APEI07Couple/common/genmapgen_ATM_T21-OCN_Pl42_conserve.conf:1-4)."""
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    A, O, _ = pair(orc, dccm, name)
    lx = X.LegacyExchange(A, O, order_ao=2 if O.im == 1 else 1, device=gpu)
    a2o, o2a = S.generic_fields(np, A, 12) + 2.0, S.generic_fields(np, O, 4, salt=5.0) + 2.0
    lx.a2o.copy_(torch.from_numpy(a2o)); lx.o2a.copy_(torch.from_numpy(o2a))
    lx.step()
    torch.cuda.synchronize()
    want_o = np.concatenate([orc.remap_apply(*lx.tabs["ao_cons"], a2o[:10], O.n), orc.remap_apply(*lx.tabs["ao_bil"], a2o[10:], O.n)])
    want_a = np.concatenate([orc.remap_apply(*lx.tabs["oa_cons"], o2a[:3], A.n), orc.remap_apply(*lx.tabs["oa_bil"], o2a[3:], A.n)])
    assert np.array_equal(lx.o_recv.cpu().numpy(), want_o) and np.array_equal(lx.a_recv.cpu().numpy(), want_a)
    # conservative A->O keeps the global integral (area weights of the destination / source rows)
    wo, wa = np.repeat(O.y_LatWt, O.im) / O.im, np.repeat(A.y_LatWt, A.im) / A.im
    assert abs((want_o[2] * wo).sum() / (a2o[2] * wa).sum() - 1.0) <= 1e-11


def test_compiled_host_driver_through_the_c_abi_equals_resident_path(gpu, dccm, tmp_path):
    """examples/exchange_driver.cpp: one exchange from compiled host code through the C ABI only (operators built from
    the grid axes, interpolate_data x8, DSFCM_Util_SfcBulkFlux_Get on (IA,JA) halo arrays, VDiffForward / Backward with
    host arrays), in the call order of the reference's component drivers.  Its dumped outputs equal, bit for bit, the
    device-resident SurfaceExchange run on the dumped inputs."""
    import subprocess
    import json
    import torch
    X = importlib.import_module("dennou-ccm_b200.exchange")
    exe = dccm.build_driver()
    ima, jma, imo, jmo, K = 64, 32, 72, 36, 8
    r = subprocess.run([exe, str(ima), str(jma), str(imo), str(jmo), str(K), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["kinds"] == [1, 1, 2, 2, 1, 1, 2, 2]          # A<->S zonal stencils, O<->S separable (bil, cons per pair)
    load = lambda name, *shape: torch.as_tensor(np.fromfile(tmp_path / f"{name}.f64").reshape(shape), device=gpu)
    T = dccm.tables
    A, O = T.get_LonLatGrid(ima, jma), T.regular_LonLatGrid(imo, jmo)
    Sx = T.generate_surface_exchange_grid(A, O)
    assert [Sx.im, Sx.jm] == info["sfc"]
    nA, nO, nS = A.n, O.n, Sx.n
    ex = X.SurfaceExchange(A, O, Sx, K, 1, 1, fast=False, device=gpu,
                           consts=dict(Grav=9.8, CpDry=1004.6, GasRDry=287.04, DelTime=1200.0, Sig1=0.995))
    col = {k: load(k, K + 1, nA) for k in ("MomFluxX", "MomFluxY", "HeatFlux", "Press", "rExner", "VirTemp",
                                            "VelDiffCoef", "TempDiffCoef", "QMixDiffCoef")}
    col["QMixFlux"] = load("QMixFlux", 1, K + 1, nA)
    col["zExner"], col["Height"] = load("zExner", K, nA), load("Height", K, nA)
    ab, ac, ob, oc = load("a2s_bil", 13, nA), load("a2s_cons", 4, nA), load("o2s_bil", 2, nO), load("o2s_cons", 3, nO)
    ex.set_inputs(col, {"WindU": ab[0:1], "WindV": ab[1:2], "SfcAirTemp": ab[2:3], "QVap1": ab[3:4], "SfcPress": ab[4:5],
                        "LDwRFlx": ac[0:1], "SDwRFlx": ac[1:2], "RainFall": ac[2:3], "SnowFall": ac[3:4]},
                  {"SfcTempO": ob[0:1], "SfcTempI": ob[1:2], "SIceCon": oc[0:1], "SfcAlbedoO": oc[1:2], "SfcAlbedoI": oc[2:3]})
    ex.step()
    torch.cuda.synchronize()
    assert torch.equal(ex.a2s_bil, ab)                        # the forward solve's coefficients, layers 5..12
    for name, got in (("s2a", ex.s2a), ("s2o", ex.s2o), ("a_recv", ex.a_recv), ("o_recv", ex.o_recv)):
        assert torch.equal(got, load(name, *got.shape)), name
    for name in ("DUDt", "DVDt", "DTempDt"):
        assert torch.equal(ex.tend[name], load(name, K, nA)), name
    assert torch.equal(ex.tend["DQMixDt"], load("DQMixDt", 1, K, nA))
    assert bool(torch.isfinite(ex.o_recv).all()) and float(ex.o_recv.abs().max()) > 0.0
