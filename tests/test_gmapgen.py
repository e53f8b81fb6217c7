"""tool/gmapgen/gmapgen_main (compiled, C ABI only) against the oracle's generators, driven by the namelist
settings of the reference's own experiment directories (ref tool/gmapgen/gmapgen_main.f90:9-149, :158-254)."""
import os
import subprocess

import numpy as np
import pytest

# namelist contents of ref exp/APEI07Couple/common/genmapgen_ATM_T21-OCN_Pl42.conf (bilinear, six files),
# exp/APEI07Couple/common/genmapgen_ATM_T42-OCN_Pl42_conserve.conf (Jones99, four files; OS / SO names not set)
# and exp/APESpinUpSolarDepExp/common/gmapgen.conf (two files, ConservativeFlag left at its default)
CONFS = {
    "T21_Pl42_bilinear": """&PARAM_DCCM_GRID
  IMA = 64, JMA = 32, NMA = 21,
  IMO =  1, JMO = 64, NMO = 42,
/
&PARAM_GMAPGEN
 gmapfile_AO_NAME = "gmap-ATM_T21-OCN_Pl42.dat",
 gmapfile_OA_NAME = "gmap-OCN_Pl42-ATM_T21.dat",
 gmapfile_SA_NAME = "gmap-SFC_T21Pl42-ATM_T21.dat",
 gmapfile_AS_NAME = "gmap-ATM_T21-SFC_T21Pl42.dat",
 gmapfile_SO_NAME = "gmap-SFC_T21Pl42-OCN_Pl42.dat",
 gmapfile_OS_NAME = "gmap-OCN_Pl42-SFC_T21Pl42.dat",
 ConservativeFlag = .false.,
/
""",
    "T42_Pl42_conserve": """&PARAM_DCCM_GRID
  IMA = 128, JMA = 64, NMA = 42,
  IMO =  1, JMO = 64, NMO = 42,
/
&PARAM_GMAPGEN
 gmapfile_AO_NAME = "gmap-ATM_T42-OCN_Pl42_conserve.dat",
 gmapfile_OA_NAME = "gmap-OCN_Pl42-ATM_T42_conserve.dat",
 gmapfile_SA_NAME = "gmap-SFC_Pl42-ATM_T42_conserve.dat",
 gmapfile_AS_NAME = "gmap-ATM_T42-SFC_Pl42_conserve.dat",
 ConservativeFlag = .true.,
/
""",
    "solar_dep_default_flag": """&PARAM_DCCM_GRID
  IMA = 64, JMA = 32, NMA = 21,
  IMO =  1, JMO = 64, NMO = 42,
/
&PARAM_GMAPGEN
 gmapfile_AO_NAME="gmap-ATM_T21-OCN_Pl42.dat",
 gmapfile_OA_NAME="gmap-OCN_Pl42-ATM_T21.dat"
/
""",
}
EXPECT = {   # conf -> (IMA, JMA, IMO, JMO, conservative, {pair: file})
    "T21_Pl42_bilinear": (64, 32, 1, 64, False, {"AO": "gmap-ATM_T21-OCN_Pl42.dat", "OA": "gmap-OCN_Pl42-ATM_T21.dat",
                                                 "SA": "gmap-SFC_T21Pl42-ATM_T21.dat", "AS": "gmap-ATM_T21-SFC_T21Pl42.dat",
                                                 "SO": "gmap-SFC_T21Pl42-OCN_Pl42.dat", "OS": "gmap-OCN_Pl42-SFC_T21Pl42.dat"}),
    "T42_Pl42_conserve": (128, 64, 1, 64, True, {"AO": "gmap-ATM_T42-OCN_Pl42_conserve.dat",
                                                 "OA": "gmap-OCN_Pl42-ATM_T42_conserve.dat",
                                                 "SA": "gmap-SFC_Pl42-ATM_T42_conserve.dat",
                                                 "AS": "gmap-ATM_T42-SFC_Pl42_conserve.dat"}),
    "solar_dep_default_flag": (64, 32, 1, 64, False, {"AO": "gmap-ATM_T21-OCN_Pl42.dat", "OA": "gmap-OCN_Pl42-ATM_T21.dat"}),
}
ORDER = {"AO": 2, "OA": 2, "AS": 2, "SA": 1, "OS": 1, "SO": 1}        # defaults, ref :214-224


@pytest.fixture(scope="module")
def gmapgen(dccm):
    return dccm.build_gmapgen()


@pytest.mark.parametrize("name", sorted(CONFS))
def test_gmapgen_writes_the_reference_files(gmapgen, orc, tmp_path, name):
    (tmp_path / "my.conf").write_text(CONFS[name])
    r = subprocess.run([gmapgen, "--N=my.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ima, jma, imo, jmo, cons, files = EXPECT[name]
    A, O = orc.gauss_grid(ima, jma), orc.gauss_grid(imo, jmo)
    G = {"A": A, "O": O, "S": orc.exchange_grid(A, O)}
    assert sorted(os.listdir(tmp_path)) == sorted(["my.conf"] + list(files.values()))
    for pair, fn in files.items():
        s, d = G[pair[0]], G[pair[1]]
        want = orc.gen_jones99(s, d, ORDER[pair]) if cons else orc.gen_bilinear(s, d)
        got = orc.read_table(str(tmp_path / fn))            # the reference's list-directed reader, restated
        for x, y in zip((got.iD, got.jD, got.iS, got.jS, got.coef), (want.iD, want.jD, want.iS, want.jS, want.coef)):
            assert np.array_equal(x, y), (pair, fn)
    assert "Set mapping table and coeffecient" in r.stdout    # check_mappingTable, ref :151-152


def test_gmapgen_defaults_orders_and_errors(gmapgen, orc, tmp_path):
    # default namelist name gmapgen.conf (ref :30); keys are case-insensitive, '!' starts a comment, single quotes work;
    # interp_order_* are honoured in conservative mode (ref :126-148)
    (tmp_path / "gmapgen.conf").write_text("""! comment line
&param_dccm_grid  ima=32, jma=16, imo=32, jmo=16 /
&PARAM_GMAPGEN
  GMAPFILE_AS_NAME = 'as.dat'   ! trailing comment with a / slash
  gmapfile_SA_name = 'sa.dat', interp_order_AS = 1, interp_order_SA = 2
  ConservativeFlag = T
/
""")
    r = subprocess.run([gmapgen, "--quiet"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout == "", r.stderr
    A = orc.gauss_grid(32, 16)
    S = orc.exchange_grid(A, A)                              # JMA == JMO: the atmosphere's rows (ref :355-360)
    for fn, (s, d, order) in {"as.dat": (A, S, 1), "sa.dat": (S, A, 2)}.items():
        got, want = orc.read_table(str(tmp_path / fn)), orc.gen_jones99(s, d, order)
        assert np.array_equal(got.coef, want.coef) and np.array_equal(got.iS, want.iS) and np.array_equal(got.jS, want.jS)
    # mismatched longitudes: the reference generator stops; so does the tool unless the extension is asked for
    (tmp_path / "mis.conf").write_text("&PARAM_DCCM_GRID IMA=32, JMA=16, IMO=24, JMO=12 /\n"
                                       "&PARAM_GMAPGEN gmapfile_AO_NAME='ao.dat', interp_order_AO=1, ConservativeFlag=.true. /\n")
    r = subprocess.run([gmapgen, "--namelist=mis.conf", "--quiet"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "lon_mode" in r.stderr
    r = subprocess.run([gmapgen, "--namelist=mis.conf", "--quiet", "--lon-mode=1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = orc.gen_jones99(orc.gauss_grid(32, 16), orc.gauss_grid(24, 12), 1, lon_mode=1)   # generalised overlap is 1st order
    assert np.array_equal(orc.read_table(str(tmp_path / "ao.dat")).coef, want.coef)
    r = subprocess.run([gmapgen, "--N=absent.conf"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open namelist" in r.stderr
