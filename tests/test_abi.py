"""CPU-side tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/dccm_b200.h declares, the host-side table code agrees index-for-index with the oracle,
and the product never routes through the oracle or a CPU fallback.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import as_orc_grid, pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dennou-ccm_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dccm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dccm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(dccm):
    lib = ctypes.CDLL(dccm._lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dccm_b200.h but not exported"
    assert sorted(dccm._lib.declared_symbols()) == syms, "ctypes binding and header disagree"
    assert b"sm_100a" in dccm.lib().dccm_build_info()


def test_library_exports_the_fortran_external_interpolate_data(dccm):
    """ref common/interpolate_data.f90:1-17, Mkinclude:167: Jcup links the bare symbol `interpolate_data_`; the library
    exports it itself.  Host-side behaviour only here: name table, blank-padded CHARACTER arguments, error paths."""
    lib = dccm.lib()
    assert hasattr(ctypes.CDLL(dccm._lib.LIB_PATH), "interpolate_data_")
    dccm._lib.f77_interpolate_data()                          # binds with the 12 + 2 hidden-length signature
    x, y = np.zeros((1, 4)), np.zeros((1, 4))
    f = dccm._lib.dp
    # unknown component name -> status + message (the Fortran entry would hand the message to the error handler)
    rc = lib.dccm_interpolate_data_named(b"SFC     ", 8, b"ATM", 3, 1, 4, 1, f(x), 4, 1, f(y), 1)
    assert rc == 1 and b"unknown component name 'SFC'" in lib.dccm_last_error()
    dccm._lib.check(lib.dccm_interp_set_model_name(3, b"SFC"))
    dccm._lib.check(lib.dccm_interp_set_model_name(1, b"ATM"))
    rc = lib.dccm_interpolate_data_named(b"SFC     ", 8, b"ATM", 3, 7, 4, 1, f(x), 4, 1, f(y), 1)
    assert rc == 1 and b"no operation index registered for (recv=3, send=1, tag=7)" in lib.dccm_last_error()
    assert lib.dccm_interp_set_model_name(0, b"X") == 1 and lib.dccm_interp_set_model_name(2, b"") == 1
    # the error handler receives the message when the Fortran-named entry fails
    seen = []
    H = ctypes.CFUNCTYPE(None, ctypes.c_char_p)
    h = H(lambda m: seen.append(m))
    lib.dccm_f77_set_error_handler(ctypes.cast(h, ctypes.c_void_p))
    try:
        i = lambda v: ctypes.byref(ctypes.c_int32(v))
        p32 = lambda v: ctypes.cast(i(v), dccm._lib.i32p)
        dccm._lib.f77_interpolate_data()(b"OCN  ", b"ATM  ", p32(1), p32(4), p32(1), f(x), p32(4), p32(1), f(y), p32(1),
                                        p32(1), p32(1), 5, 5)
        assert seen and b"unknown component name 'OCN'" in seen[0]
    finally:
        lib.dccm_f77_set_error_handler(None)


def test_library_is_sm100a_only(dccm):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", dccm._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_does_not_touch_the_oracle():
    for dp, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.lower().replace("bit-exact vs the oracle", "").replace(
                    "bit-identical to the reference", ""), f"{f} mentions the oracle"


@pytest.mark.parametrize("name", ["T21_Pl42", "T42_T42", "T21_1deg", "T106_1deg"])
def test_generators_match_oracle_index_for_index(orc, dccm, name):
    """'identical remap indices': same entries, same order, bit-identical weights."""
    T = dccm.tables
    A, O, S = pair(orc, dccm, name)
    oA, oO, oS = [as_orc_grid(orc, g) for g in (A, O, S)]
    for (s, d, os_, od) in [(A, S, oA, oS), (S, A, oS, oA), (S, O, oS, oO), (O, S, oO, oS)]:
        for order in (1, 2):
            try:
                t = T.gen_table_jones99(s, d, order, 1).entries()
            except dccm.DccmError:
                with pytest.raises(RuntimeError):
                    orc.gen_jones99(os_, od, order, 1)
                continue
            o = orc.gen_jones99(os_, od, order, 1)
            for x, y in zip(t, (o.iD, o.jD, o.iS, o.jS, o.coef)):
                assert np.array_equal(x, y)
        for lm in (0, 1):
            t = T.gen_table_bilinear(s, d, lm).entries()
            o = orc.gen_bilinear(os_, od, lm)
            for x, y in zip(t, (o.iD, o.jD, o.iS, o.jS, o.coef)):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("sizes", [(8, 5, 4, 3), (360, 181, 128, 65), (128, 65, 360, 181), (7, 4, 7, 4),
                                   (1, 2, 5, 3), (5, 3, 1, 2), (96, 49, 100, 37)])
def test_make_mapping_table_matches_oracle_entry_for_entry(orc, dccm, sizes, tmp_path):
    """ref common/cal_mappingtable.f90:10-49: same entries, same order, bit-identical weights; and the file the
    mirror module writes reads back to the same 1-D indices."""
    o = orc.make_mapping_table(*sizes)
    t = dccm.tables.make_mapping_table(*sizes)
    for x, y in zip(t.entries(), (o.iD, o.jD, o.iS, o.jS, o.coef)):
        assert np.array_equal(x, y)
    f = str(tmp_path / "mapping_table.txt")
    dccm.cal_mappingtable.make_mapping_table(f, *sizes)
    back = orc.read_table(f)
    for x, y in zip(o.to_index(sizes[2], sizes[0]), back.to_index(sizes[2], sizes[0])):
        assert np.array_equal(x, y)


def test_make_mapping_table_rejects_degenerate_grids(dccm):
    with pytest.raises(dccm.DccmError, match="ny >= 2"):
        dccm.tables.make_mapping_table(8, 1, 4, 3)


def test_unsupported_grid_pair_fails_like_the_reference(dccm, orc):
    A, O, S = pair(orc, dccm, "T21_1deg")
    with pytest.raises(dccm.DccmError, match="lon_mode=1"):
        dccm.tables.gen_table_jones99(O, S, 1, 0)
    with pytest.raises(dccm.DccmError, match="2nd order"):
        dccm.tables.gen_table_jones99(O, S, 2, 1)


def test_reference_named_table_interfaces(dccm, orc, tmp_path):
    """gen_gridmapfile_lonlat2lonlat + set_mappingTable_interpCoef with the reference's argument
    lists (ref common/grid_mapping_util_jones99.f90:35-54,446-462; common/grid_mapping_util.f90:32-48)."""
    A, O, S = pair(orc, dccm, "T21_Pl42")
    fn = str(tmp_path / "gmap-ATM_T21-SFC_conserve.dat")
    dccm.grid_mapping_util_jones99.gen_gridmapfile_lonlat2lonlat(
        fn, A.x_Lon, A.y_Lat, S.x_Lon, S.y_Lat, A.x_LonWt, A.y_LatWt, S.x_LonWt, S.y_LatWt, 2)
    send, recv, coef = dccm.grid_mapping_util_jones99.set_mappingTable_interpCoef(fn, A.im, S.im)
    o = orc.gen_jones99(as_orc_grid(orc, A), as_orc_grid(orc, S), 2)
    os_, or_, oc = o.to_index(A.im, S.im)
    assert np.array_equal(send, os_) and np.array_equal(recv, or_) and np.array_equal(coef, oc)
    # the oracle reads the product's file and vice versa (same on-disk format)
    r = orc.read_table(fn)
    assert np.array_equal(r.coef, o.coef) and np.array_equal(r.jS, o.jS)
    fn2 = str(tmp_path / "gmap-bilinear.dat")
    dccm.grid_mapping_util.gen_gridmapfile_lonlat2lonlat(fn2, S.x_Lon, S.y_Lat, O.x_Lon, O.y_Lat)
    send, recv, coef = dccm.grid_mapping_util.set_mappingTable_interpCoef(fn2, S.im, O.im)
    ob = orc.gen_bilinear(as_orc_grid(orc, S), as_orc_grid(orc, O))
    os_, or_, oc = ob.to_index(S.im, O.im)
    assert np.array_equal(send, os_) and np.array_equal(recv, or_) and np.array_equal(coef, oc)
    # binary table form round-trips exactly
    t = dccm.tables.gen_table_jones99(A, S, 2)
    t.write(str(tmp_path / "t.bin"), binary=True)
    u = dccm.tables.MappingTable.read(str(tmp_path / "t.bin"), binary=True)
    for x, y in zip(t.entries(), u.entries()):
        assert np.array_equal(x, y)
    with pytest.raises(dccm.DccmError):
        dccm.tables.MappingTable.read(str(tmp_path / "missing.dat"))


def test_dsfcm_admin_grid_layout(dccm):
    """ref sfc/DSFCM_Admin_Grid_mod.f90:35-52 and sfc/DSFCM_Admin_Variable_mod.f90:57-80"""
    g = dccm.dsfcm.DSFCM_Admin_Grid(128, 64)
    assert (g.IS, g.IE, g.IA, g.JS, g.JE, g.JA) == (2, 129, 130, 2, 65, 66)
    v = dccm.dsfcm.DSFCM_Admin_Variable(g)
    assert v.xya_SenHFlx.shape == (3, 66, 130) and v.xya_DelVarImplCPL.shape == (4, 66, 130)
    assert v.xy_SIceCon.shape == (66, 130)


def test_header_is_plain_c_and_links_from_c(dccm, tmp_path):
    """The boundary is a C ABI: include/dccm_b200.h compiles as strict C99 and a C client links against the library
    (host-only entry points here; dccm_init must fail loudly without a GPU instead of falling back)."""
    import subprocess
    src = tmp_path / "client.c"
    src.write_text(r"""
#include "dccm_b200.h"
#include <stdio.h>
int main(void) {
    double lon[8], lat[4], lw[8], aw[4];
    dccm_table *t = 0;
    if (dccm_grid_gauss(8, 4, lon, lat, lw, aw)) return 1;
    if (dccm_table_gen_make_mapping_table(8, 5, 4, 3, &t)) return 2;
    printf("entries=%lld\n", (long long)dccm_table_size(t));
    dccm_table_free(t);
    if (dccm_table_gen_make_mapping_table(8, 1, 4, 3, &t) != DCCM_ERR_ARG) return 3;
    printf("error=%s\n", dccm_last_error());
    return 0;
}
""")
    exe = tmp_path / "client"
    libdir = os.path.dirname(dccm._lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-o", str(exe), str(src), "-L", libdir, "-ldccm_b200", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "entries=84" in r.stdout and "ny >= 2" in r.stdout, r.stdout + r.stderr


def test_staged_surface_kernel_is_compiled_with_tma_bulk_copies(dccm):
    """B200_PROFILING.md: the SASS mnemonics that prove the TMA path -- UBLKCP (cp.async.bulk global -> shared) and
    SYNCS (mbarrier arrive / try_wait) -- must appear in the staged fused surface kernel and nowhere tensor-core
    instructions are expected (the path is HBM-bound fp64: no HMMA / UTCMMA in the library)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", dccm._lib.LIB_PATH], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    staged = [f for f in funcs if "sfc_exchange_staged_kernel" in f.split("\n", 1)[0]]
    assert len(staged) >= 6                                   # MINB 4/5/6 x peer-segment on/off (+ the API-complete form)
    for f in staged:
        assert "UBLKCP" in f and "SYNCS" in f, f.split("\n", 1)[0]
    assert not re.search(r"\b(HMMA|UTCHMMA|UTCMMA|IMMA|DMMA)\b", sass)
    others = [f for f in funcs if "UBLKCP" in f and "sfc_exchange_staged_kernel" not in f.split("\n", 1)[0]]
    assert not others


def test_no_cpu_fallback_without_gpu(dccm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dccm.DccmError, match="no CPU fallback"):
        dccm.RemapOperator([1], [1], [1.0], 1, 1)
    with pytest.raises(dccm.DccmError, match="no CPU fallback"):
        dccm.SfcImplicitCoupling(4, 2, 5, 1, 1, 9.8, 1004.6, 287.04, 1200.0)
    z = np.zeros((3, 5, 6))
    with pytest.raises(dccm.DccmError, match="no CPU fallback"):
        dccm.DSFCM_Util_SfcBulkFlux_Get(6, 5, *[z.copy() for _ in range(8)], np.zeros((4, 5, 6)),
                                        *[z.copy() for _ in range(5)], *[np.zeros((5, 6)) for _ in range(6)],
                                        np.zeros((4, 5, 6)), np.zeros((4, 5, 6)), z.copy(), z.copy(),
                                        np.zeros((5, 6)), np.array([0.995, 0.01]), np.zeros((5, 6)), np.zeros((5, 6)))


def test_zonal_stencil_detection_is_exact(dccm, orc):
    """Every table the reference generator can produce (equal longitudes, or an axisymmetric source)
    repeats one stencil along each destination latitude row; the library then keeps O(ny) entries
    instead of O(nx*ny).  Detection is bitwise and host-only (dccm_remap_classify)."""
    import ctypes as C
    L = dccm._lib

    def classify(tab, src, dst, hint=True):
        s, r, c = tab.index(src.im, dst.im)
        k, n = C.c_int(), C.c_int64()
        L.check(L.lib().dccm_remap_classify(len(c), L.ip(s), L.ip(r), L.dp(c), src.n, dst.n,
                                            src.im if hint else 0, dst.im if hint else 0, C.byref(k), C.byref(n)))
        return k.value, n.value, len(c)

    T = dccm.tables
    A, O, S = pair(orc, dccm, "T106_1deg")
    for tab, s, d in ((T.gen_table_jones99(A, S, 2), A, S), (T.gen_table_bilinear(A, S), A, S),
                      (T.gen_table_jones99(S, A, 1), S, A), (T.gen_table_bilinear(S, A), S, A)):
        kind, kept, nnz = classify(tab, s, d)
        assert kind == 1 and kept * d.im == nnz
        assert classify(tab, s, d, hint=False)[0] == 0          # no grid hint -> general CSR
    # mismatched longitudes (320 vs 360): not shift invariant -> CSR
    assert classify(T.gen_table_jones99(O, S, 1, 1), O, S)[0] == 0
    assert classify(T.gen_table_bilinear(S, O, 1), S, O)[0] == 0
    # axisymmetric source (shipped Pl42 ocean): zonal; axisymmetric destination: one long row -> CSR
    A, O, S = pair(orc, dccm, "T21_Pl42")
    assert classify(T.gen_table_jones99(O, S, 1), O, S)[0] == 1
    assert classify(T.gen_table_jones99(S, O, 1), S, O)[0] == 0
    # a single perturbed weight breaks the pattern and must be noticed
    s, r, c = T.gen_table_jones99(S, A, 1).index(S.im, A.im)
    c2 = c.copy(); c2[len(c2) // 2] = np.nextafter(c2[len(c2) // 2], 2.0)
    k, n = C.c_int(), C.c_int64()
    L.check(L.lib().dccm_remap_classify(len(c2), L.ip(s), L.ip(r), L.dp(c2), S.n, A.n, S.im, A.im, C.byref(k), C.byref(n)))
    assert k.value == 0


@pytest.mark.parametrize("name", ["T21_1deg", "T106_1deg"])
def test_separable_factors_multiply_out_to_the_generated_tables(dccm, name):
    """for grid pairs with different longitudes the kind-2 operator keeps per-column longitude
    factors and per-row latitude factors only; multiplied out in the kernels' order (with the generator's 1e-14
    drop test) they give the generators' tables entry for entry, bit for bit.  Equal longitudes are refused
    (those tables are zonal stencils)."""
    from util import pair
    T = dccm.tables
    A, O, S = pair(None, dccm, name)
    for src, dst in ((O, S), (S, O), (A, O), (O, A)):
        for cons in (True, False):
            want = (T.gen_table_jones99(src, dst, 1, 1) if cons else T.gen_table_bilinear(src, dst, 1)).entries()
            got = T.gen_table_separable(src, dst, cons).entries()
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (name, src.im, dst.im, cons)
    # equal longitudes: described directly as one stencil per destination row (kind 1), multiplied out the same way
    for src, dst in ((A, S), (S, A)):
        for order in (1, 2):
            for a, b in zip(T.gen_table_separable(src, dst, True, order).entries(), T.gen_table_jones99(src, dst, order, 1).entries()):
                assert np.array_equal(a, b), (name, order)
        for a, b in zip(T.gen_table_separable(src, dst, False).entries(), T.gen_table_bilinear(src, dst, 1).entries()):
            assert np.array_equal(a, b), name


def test_zonal_description_of_axisymmetric_pairs(dccm):
    """nx == 1 on either side (the shipped APEI07Couple ocean): the conservative generator's table is one stencil per
    destination row; its direct description multiplies out to the generated table."""
    from util import pair
    T = dccm.tables
    A, O, S = pair(None, dccm, "T21_Pl42")
    assert O.im == 1
    for src, dst in ((O, S), (S, O), (A, O), (O, A)):
        for a, b in zip(T.gen_table_separable(src, dst, True).entries(), T.gen_table_jones99(src, dst, 1, 0).entries()):
            assert np.array_equal(a, b), (src.im, dst.im)
    with pytest.raises(dccm.DccmError):              # bilinear with an axisymmetric side: through the table only
        T.gen_table_separable(O, S, False)
