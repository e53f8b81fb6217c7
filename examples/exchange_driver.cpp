// exchange_driver.cpp -- one coupling exchange driven from compiled host code through the C ABI only.
//
// What a maintainer gets after linking fortran/*.f90: the call sequence of the reference's component drivers
// (atm: dcpam_Prepair_ImplcitCoupling -> jcup put, ref atm/dccm_atm_mod.f90:697-712; sfc: get -> DSFCM bulk flux
// -> put, ref sfc/dccm_sfc_mod.f90:865-897, :764-809; atm: get -> level-1 update -> VDiffBackward,
// ref atm/dccm_atm_mod.f90:817-835; ocn: get, ref ocn/dccm_ocn_mod.f90:978-993) with HOST arrays in and out of
// every call and the unchanged interfaces (interpolate_data, DSFCM_Util_SfcBulkFlux_Get,
// SfcImplicitCoupling_VDiffForward / Backward).  No Python, no torch: include/dccm_b200.h + libdccm_b200.so.
//
//   exchange_driver IMA JMA IMO JMO KMAX [dump_dir]
//
// Grids: Gaussian IMA x JMA atmosphere, regular IMO x JMO ocean, merged exchange grid.  Inputs are smooth synthetic
// fields (closed-form, reproducible).  With dump_dir every input and output array is written as raw float64
// (tests/test_gpu_parity.py re-runs the same inputs through the device-resident path and compares bit for bit).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/dccm_b200.h"

namespace {

using vec = std::vector<double>;

void check(int rc, const char *where)
{
    if (rc == 0) return;
    std::fprintf(stderr, "%s: libdccm_b200 error %d: %s\n", where, rc, dccm_last_error());
    std::exit(1);
}

struct Grid {
    int im = 0, jm = 0;
    vec lon, lat, lonwt, latwt;
    int n() const { return im * jm; }
};

Grid make(int im, int jm, bool gauss)
{
    Grid g;
    g.im = im; g.jm = jm;
    g.lon.resize(im); g.lat.resize(jm); g.lonwt.resize(im); g.latwt.resize(jm);
    check((gauss ? dccm_grid_gauss : dccm_grid_regular)(im, jm, g.lon.data(), g.lat.data(), g.lonwt.data(), g.latwt.data()),
          "grid");
    return g;
}

dccm_remap *op(const Grid &s, const Grid &d, bool cons)
{
    dccm_remap *h = nullptr;
    if (cons)
        check(dccm_remap_create_jones99(s.im, s.lon.data(), s.jm, s.lat.data(), d.im, d.lon.data(), d.jm, d.lat.data(),
                                        s.latwt.data(), d.latwt.data(), 1, 1, &h), "remap_create_jones99");
    else
        check(dccm_remap_create_bilinear(s.im, s.lon.data(), s.jm, s.lat.data(), d.im, d.lon.data(), d.jm, d.lat.data(), 1, &h),
              "remap_create_bilinear");
    return h;
}

// smooth, positive where physics needs it; c = cell index on the grid
double wave(const Grid &g, int c, double kx, double ky, double phase)
{
    const int j = c / g.im, i = c - j * g.im;
    return std::sin(kx * g.lon[i] + phase) * std::cos(ky * g.lat[j]);
}

void dump(const std::string &dir, const char *name, const vec &v)
{
    if (dir.empty()) return;
    const std::string path = dir + "/" + name + ".f64";
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(v.data(), 8, v.size(), f) != v.size()) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); std::exit(1); }
    std::fclose(f);
}

double sum(const vec &v) { double s = 0.0; for (double x : v) s += x; return s; }

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: %s IMA JMA IMO JMO KMAX [dump_dir]\n", argv[0]); return 2; }
    const int ima = std::atoi(argv[1]), jma = std::atoi(argv[2]), imo = std::atoi(argv[3]), jmo = std::atoi(argv[4]);
    const int K = std::atoi(argv[5]);
    const std::string dir = argc > 6 ? argv[6] : "";
    const int ATM = 1, OCN = 2, SFC = 3, BIL = 1, CONS = 2;       // Jcup component numbers / mapping tags
    check(dccm_init(0), "dccm_init");

    // ---- grids and operators (what gmapgen + set_mappingTable_interpCoef + set_interpolate_coef do at init)
    Grid A = make(ima, jma, true), O = make(imo, jmo, false), S;
    S.im = ima; S.lon = A.lon; S.lonwt = A.lonwt;
    S.lat.resize(jma + jmo); S.latwt.resize(jma + jmo);
    check(dccm_grid_exchange(jma, A.lat.data(), A.latwt.data(), jmo, O.latwt.data(), &S.jm, S.lat.data(), S.latwt.data()), "grid_exchange");
    S.lat.resize(S.jm); S.latwt.resize(S.jm);
    struct { int recv, send; const Grid *s, *d; } pairs[] = {{SFC, ATM, &A, &S}, {SFC, OCN, &O, &S}, {ATM, SFC, &S, &A}, {OCN, SFC, &S, &O}};
    std::vector<dccm_remap *> ops;
    for (auto &p : pairs)
        for (int tag : {BIL, CONS}) {
            ops.push_back(op(*p.s, *p.d, tag == CONS));
            check(dccm_interp_register(p.recv, p.send, tag, ops.back()), "interp_register");
        }
    const int nA = A.n(), nS = S.n(), nO = O.n(), nc = 1;
    const double Grav = 9.8, CpDry = 1004.6, GasRDry = 287.04, DelTime = 1200.0, sig1 = 0.995;
    dccm_vdiff *vd = nullptr;
    check(dccm_vdiff_create(ima, jma, K, nc, 1, Grav, CpDry, GasRDry, DelTime, &vd), "vdiff_create");

    // ---- synthetic column inputs, (level, column) with the column index fastest
    auto half = [&](auto f) { vec v((size_t)(K + 1) * nA); for (int l = 0; l <= K; l++) for (int c = 0; c < nA; c++) v[(size_t)l * nA + c] = f(l, c); return v; };
    auto full = [&](auto f) { vec v((size_t)K * nA); for (int k = 1; k <= K; k++) for (int c = 0; c < nA; c++) v[(size_t)(k - 1) * nA + c] = f(k, c); return v; };
    auto sigh = [&](int l) { return std::pow(1.0 - (double)l / K, 1.5); };                   // half-level sigma, 1 .. 0
    vec Press = half([&](int l, int c) { return (1.0e5 + 500.0 * wave(A, c, 2, 1, 0.3)) * sigh(l) + 10.0; });
    vec rExner = half([&](int l, int c) { return std::pow(Press[(size_t)l * nA + c] / 1.0e5, GasRDry / CpDry); });
    vec zExner = full([&](int k, int c) { return 0.5 * (rExner[(size_t)(k - 1) * nA + c] + rExner[(size_t)k * nA + c]); });
    vec VirTemp = half([&](int l, int c) { return 288.0 - 60.0 * (1.0 - sigh(l)) + 2.0 * wave(A, c, 3, 2, 1.0); });
    vec Height = full([&](int k, int c) { return 50.0 + 900.0 * (k - 1) * (1.0 + 0.05 * k) + 5.0 * wave(A, c, 1, 1, 0.0); });
    auto diff = [&](double amp) { return half([&, amp](int l, int c) { return (l == 0 || l == K) ? 0.0 : amp * std::exp(-0.25 * l) + 0.1 + 0.01 * wave(A, c, 2, 2, 0.5); }); };
    vec DiffV = diff(10.0), DiffT = diff(12.0), DiffQ = diff(11.0);
    vec FX = half([&](int l, int c) { return 0.1 * std::exp(-0.3 * l) * (1.0 + wave(A, c, 1, 1, 0.2)); });
    vec FY = half([&](int l, int c) { return -0.05 * std::exp(-0.3 * l) * (1.0 + wave(A, c, 2, 1, 0.7)); });
    vec FH = half([&](int l, int c) { return 20.0 * std::exp(-0.4 * l) * (1.0 + 0.5 * wave(A, c, 1, 2, 1.1)); });
    vec FQ = half([&](int l, int c) { return 3.0e-5 * std::exp(-0.5 * l) * (1.0 + 0.5 * wave(A, c, 2, 2, 0.4)); });
    vec DU((size_t)K * nA), DV(DU), DT(DU), DQ(DU), Coef1((size_t)4 * nA), Coef2(Coef1);

    // ---- atmosphere: forward solve, then put (17 layers for the surface component)
    check(dccm_vdiff_forward_host(vd, FX.data(), FY.data(), FH.data(), FQ.data(), Press.data(), zExner.data(), rExner.data(),
                                  VirTemp.data(), Height.data(), DiffV.data(), DiffT.data(), DiffQ.data(),
                                  DU.data(), DV.data(), DT.data(), DQ.data(), Coef1.data(), Coef2.data()), "vdiff_forward");
    vec a2s_bil((size_t)13 * nA), a2s_cons((size_t)4 * nA), o2s_bil((size_t)2 * nO), o2s_cons((size_t)3 * nO);
    for (int c = 0; c < nA; c++) {
        const int j = c / A.im;
        const double cl = std::cos(A.lat[j]);
        a2s_bil[c] = 8.0 * wave(A, c, 1, 1, 0.0);                                            // WindU
        a2s_bil[(size_t)nA + c] = 6.0 * wave(A, c, 2, 1, 0.9);                                // WindV
        a2s_bil[(size_t)2 * nA + c] = 288.0 - 40.0 * (1.0 - cl * cl) + 2.0 * wave(A, c, 3, 1, 0.1);   // SfcAirTemp
        a2s_bil[(size_t)4 * nA + c] = 1.0e5 + 500.0 * wave(A, c, 2, 1, 0.3);                  // SfcPress
        const double T = a2s_bil[(size_t)2 * nA + c], P = a2s_bil[(size_t)4 * nA + c];
        a2s_bil[(size_t)3 * nA + c] = 0.8 * 611.0 / P * std::exp(2425300.0 / (8.3144621 / 0.018) * (1.0 / 273.0 - 1.0 / T));   // QVap1
        a2s_cons[c] = 300.0 + 20.0 * wave(A, c, 1, 2, 0.6);                                  // LDwRFlx
        a2s_cons[(size_t)nA + c] = 340.0 * cl * (0.75 + 0.25 * wave(A, c, 2, 1, 0.2));        // SDwRFlx
        a2s_cons[(size_t)2 * nA + c] = 3.0e-5 * (1.0 + wave(A, c, 1, 1, 1.3));                // RainFall
        a2s_cons[(size_t)3 * nA + c] = 1.0e-5 * (1.0 + wave(A, c, 2, 2, 2.1));                // SnowFall
    }
    for (int s = 0; s < 4; s++)
        for (int c = 0; c < nA; c++) {
            a2s_bil[(size_t)(5 + s) * nA + c] = Coef1[(size_t)s * nA + c];
            a2s_bil[(size_t)(9 + s) * nA + c] = Coef2[(size_t)s * nA + c];
        }
    for (int c = 0; c < nO; c++) {
        const int j = c / O.im;
        const double deg = std::fabs(O.lat[j]) * 180.0 / std::acos(-1.0);
        const double ice = std::fmin(std::fmax((deg - 60.0) / 20.0, 0.0), 1.0);
        const double ts = std::fmax(271.35, 288.0 - 40.0 * std::pow(std::sin(O.lat[j]), 2) + 1.5 * wave(O, c, 2, 1, 0.8));
        o2s_bil[c] = ts;                                      // SfcTemp (ocean)
        o2s_bil[(size_t)nO + c] = std::fmin(273.15, ts - 5.0);  // SfcTemp (sea ice)
        o2s_cons[c] = ice; o2s_cons[(size_t)nO + c] = 0.1; o2s_cons[(size_t)2 * nO + c] = 0.6;
    }

    // ---- surface component: get (interpolate_data x4), bulk flux on the (IA,JA) halo arrays, put
    vec s_bil((size_t)13 * nS), s_cons((size_t)4 * nS), s_obil((size_t)2 * nS), s_ocons((size_t)3 * nS);
    check(dccm_interpolate_data(SFC, ATM, BIL, nA, 13, a2s_bil.data(), nS, 13, s_bil.data(), 13), "interpolate_data A->S bil");
    check(dccm_interpolate_data(SFC, ATM, CONS, nA, 4, a2s_cons.data(), nS, 4, s_cons.data(), 4), "interpolate_data A->S cons");
    check(dccm_interpolate_data(SFC, OCN, BIL, nO, 2, o2s_bil.data(), nS, 2, s_obil.data(), 2), "interpolate_data O->S bil");
    check(dccm_interpolate_data(SFC, OCN, CONS, nO, 3, o2s_cons.data(), nS, 3, s_ocons.data(), 3), "interpolate_data O->S cons");
    const int IA = S.im + 2, JA = S.jm + 2;
    const size_t N2 = (size_t)IA * JA;
    auto halo = [&](int slots) { return vec(N2 * slots, 1.0); };
    auto unpack = [&](vec &dst, int slot, const double *src) {               // ref sfc/dccm_sfc_mod.f90:900-951
        for (int j = 0; j < S.jm; j++)
            for (int i = 0; i < S.im; i++) dst[slot * N2 + (size_t)(j + 1) * IA + i + 1] = src[(size_t)j * S.im + i];
    };
    auto pack = [&](double *dst, const vec &src, int slot, double sign) {     // ref :787-809
        for (int j = 0; j < S.jm; j++)
            for (int i = 0; i < S.im; i++) dst[(size_t)j * S.im + i] = sign * src[slot * N2 + (size_t)(j + 1) * IA + i + 1];
    };
    vec WindU = halo(1), WindV = halo(1), T1 = halo(1), Q1 = halo(1), Ps = halo(1), SDw = halo(1), LDw = halo(1), Ice = halo(1);
    vec Hgt(N2, 0.0), C1 = halo(4), C2 = halo(4), Ts = halo(3), Al = halo(3);
    unpack(WindU, 0, &s_bil[0]); unpack(WindV, 0, &s_bil[(size_t)nS]); unpack(T1, 0, &s_bil[(size_t)2 * nS]);
    unpack(Q1, 0, &s_bil[(size_t)3 * nS]); unpack(Ps, 0, &s_bil[(size_t)4 * nS]);
    for (int s = 0; s < 4; s++) { unpack(C1, s, &s_bil[(size_t)(5 + s) * nS]); unpack(C2, s, &s_bil[(size_t)(9 + s) * nS]); }
    unpack(LDw, 0, &s_cons[0]); unpack(SDw, 0, &s_cons[(size_t)nS]);
    unpack(Ts, 0, &s_obil[0]); unpack(Ts, 1, &s_obil[(size_t)nS]);
    unpack(Ice, 0, &s_ocons[0]); unpack(Al, 0, &s_ocons[(size_t)nS]); unpack(Al, 1, &s_ocons[(size_t)2 * nS]);
    vec WSX = halo(3), WSY = halo(3), SenH = halo(3), QVapM = halo(3), LatH = halo(3), VelTC = halo(3), TempTC = halo(3),
        QVapTC = halo(3), Del = halo(4), SUw = halo(3), LUw = halo(3), HFns = halo(3), HFsr = halo(3), DHF = halo(3);
    const double sig[2] = {sig1, 0.01};
    check(dccm_bulkflux_get_host(IA, JA, WSX.data(), WSY.data(), SenH.data(), QVapM.data(), LatH.data(), VelTC.data(),
                                 TempTC.data(), QVapTC.data(), Del.data(), SUw.data(), LUw.data(), HFns.data(), HFsr.data(),
                                 DHF.data(), WindU.data(), WindV.data(), T1.data(), Q1.data(), SDw.data(), LDw.data(),
                                 C1.data(), C2.data(), Ts.data(), Al.data(), Ice.data(), sig, Hgt.data(), Ps.data()), "bulkflux_get");
    vec s2a((size_t)9 * nS), s2o((size_t)12 * nS);                            // put-side selection, ref :764-784
    pack(&s2a[0], LUw, 2, 1.0); pack(&s2a[(size_t)nS], SUw, 2, 1.0); pack(&s2a[(size_t)2 * nS], SenH, 2, 1.0);
    pack(&s2a[(size_t)3 * nS], QVapM, 2, 1.0); pack(&s2a[(size_t)4 * nS], Al, 2, 1.0);
    for (int s = 0; s < 4; s++) pack(&s2a[(size_t)(5 + s) * nS], Del, s, 1.0);
    pack(&s2o[0], HFns, 0, 1.0); pack(&s2o[(size_t)nS], HFsr, 0, 1.0);
    for (int c = 0; c < nS; c++) { s2o[(size_t)2 * nS + c] = s_cons[(size_t)3 * nS + c]; s2o[(size_t)3 * nS + c] = s_cons[(size_t)2 * nS + c]; }
    pack(&s2o[(size_t)4 * nS], QVapM, 0, 1.0); pack(&s2o[(size_t)5 * nS], WSX, 2, -1.0); pack(&s2o[(size_t)6 * nS], WSY, 2, -1.0);
    pack(&s2o[(size_t)7 * nS], HFns, 1, 1.0); pack(&s2o[(size_t)8 * nS], HFsr, 1, 1.0); pack(&s2o[(size_t)9 * nS], QVapM, 1, 1.0);
    pack(&s2o[(size_t)10 * nS], DHF, 0, 1.0); pack(&s2o[(size_t)11 * nS], DHF, 1, 1.0);

    // ---- atmosphere and ocean: get, level-1 update, backward solve
    vec a_recv((size_t)9 * nA), o_recv((size_t)12 * nO);
    check(dccm_interpolate_data(ATM, SFC, CONS, nS, 4, &s2a[0], nA, 4, &a_recv[0], 4), "interpolate_data S->A cons");
    check(dccm_interpolate_data(ATM, SFC, BIL, nS, 5, &s2a[(size_t)4 * nS], nA, 5, &a_recv[(size_t)4 * nA], 5), "interpolate_data S->A bil");
    check(dccm_interpolate_data(OCN, SFC, CONS, nS, 10, &s2o[0], nO, 10, &o_recv[0], 10), "interpolate_data S->O cons");
    check(dccm_interpolate_data(OCN, SFC, BIL, nS, 2, &s2o[(size_t)10 * nS], nO, 2, &o_recv[(size_t)10 * nO], 2), "interpolate_data S->O bil");
    for (int c = 0; c < nA; c++) {                                            // ref atm/dccm_atm_mod.f90:832-835
        DU[c] = a_recv[(size_t)5 * nA + c]; DV[c] = a_recv[(size_t)6 * nA + c];
        DT[c] = a_recv[(size_t)7 * nA + c]; DQ[c] = a_recv[(size_t)8 * nA + c];
    }
    check(dccm_vdiff_backward_host(vd, DU.data(), DV.data(), DT.data(), DQ.data()), "vdiff_backward");

    const struct { const char *name; const vec *v; } out[] = {
        {"MomFluxX", &FX}, {"MomFluxY", &FY}, {"HeatFlux", &FH}, {"QMixFlux", &FQ}, {"Press", &Press}, {"zExner", &zExner},
        {"rExner", &rExner}, {"VirTemp", &VirTemp}, {"Height", &Height}, {"VelDiffCoef", &DiffV}, {"TempDiffCoef", &DiffT},
        {"QMixDiffCoef", &DiffQ}, {"a2s_bil", &a2s_bil}, {"a2s_cons", &a2s_cons}, {"o2s_bil", &o2s_bil}, {"o2s_cons", &o2s_cons},
        {"s2a", &s2a}, {"s2o", &s2o}, {"a_recv", &a_recv}, {"o_recv", &o_recv}, {"DUDt", &DU}, {"DVDt", &DV}, {"DTempDt", &DT},
        {"DQMixDt", &DQ}};
    for (auto &o : out) dump(dir, o.name, *o.v);
    std::printf("{\"atm\": [%d, %d], \"ocn\": [%d, %d], \"sfc\": [%d, %d], \"kmax\": %d, \"kinds\": [", ima, jma, imo, jmo, S.im, S.jm, K);
    for (size_t k = 0; k < ops.size(); k++) std::printf("%s%d", k ? ", " : "", dccm_remap_kind(ops[k]));
    std::printf("], \"sum_a_recv\": %.17g, \"sum_o_recv\": %.17g, \"sum_DTempDt\": %.17g}\n", sum(a_recv), sum(o_recv), sum(DT));
    for (auto h : ops) dccm_remap_destroy(h);
    dccm_vdiff_destroy(vd);
    return 0;
}
