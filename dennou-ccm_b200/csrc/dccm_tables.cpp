// dccm_tables.cpp -- host side of the mapping tables: grid axes, Jones (1999) and bilinear
// generators, exchange-grid construction, text/binary table files.
//
// One-time (init) work: the hot path only consumes the finished tables.  The generators keep
// the reference's floating-point expressions and emission order so that indices are identical
// and weights bit-identical to the reference algorithm, but the searches are hoisted: the
// latitude overlap range depends only on the destination row and the longitude range only on
// the destination column, so they are found once per row / column instead of once per cell
// (the reference rescans both for every cell, common/grid_mapping_util_jones99.f90:230-235).
#include <cmath>
#include <algorithm>

#include "dccm_common.h"
#include "dccm_sep.h"

struct dccm_table {
    std::vector<int32_t> iD, jD, iS, jS;
    std::vector<double> coef;
    void push(int a, int b, int c, int d, double w)
    {
        iD.push_back(a); jD.push_back(b); iS.push_back(c); jS.push_back(d); coef.push_back(w);
    }
    void reserve(size_t n)
    {
        iD.reserve(n); jD.reserve(n); iS.reserve(n); jS.reserve(n); coef.reserve(n);
    }
};

using namespace dccm;

namespace {

const double kPi = std::acos(-1.0);

// Gauss-Legendre nodes/weights by Newton iteration on P_n (ascending nodes, sum(w) = 2).
void gauss_legendre(int n, std::vector<double> &mu, std::vector<double> &w)
{
    mu.assign(n, 0.0);
    w.assign(n, 0.0);
    auto eval = [n](double x, double &pn, double &pnm1) {
        double p0 = 1.0, p1 = x;
        for (int l = 2; l <= n; l++) {
            double p2 = ((2.0 * l - 1.0) * x * p1 - (l - 1.0) * p0) / l;
            p0 = p1;
            p1 = p2;
        }
        pn = p1;
        pnm1 = p0;
    };
    for (int i = 0; i < (n + 1) / 2; i++) {
        double x = std::cos(kPi * (i + 0.75) / (n + 0.5));
        double pn, pm, dp = 1.0;
        for (int it = 0; it < 100; it++) {
            eval(x, pn, pm);
            dp = n * (x * pn - pm) / (x * x - 1.0);
            double dx = pn / dp;
            x -= dx;
            if (std::fabs(dx) < 1e-16) break;
        }
        eval(x, pn, pm);
        dp = n * (x * pn - pm) / (x * x - 1.0);
        double wi = 2.0 / ((1.0 - x * x) * dp * dp);
        mu[i] = -x; mu[n - 1 - i] = x;
        w[i] = wi;  w[n - 1 - i] = wi;
    }
    if (n % 2 == 1) mu[n / 2] = 0.0;
}

// cell edges, ref common/grid_mapping_util_jones99.f90:88-97 (halo), :142-151
std::vector<double> lon_edges(int n, const double *x)
{
    std::vector<double> xc(n + 2), u(n + 1);
    for (int i = 0; i < n; i++) xc[i + 1] = x[i];
    xc[0] = xc[n] - 2.0 * kPi;
    xc[n + 1] = xc[1] + 2.0 * kPi;
    for (int i = 0; i <= n; i++) u[i] = 0.5 * (xc[i] + xc[i + 1]);
    return u;
}

std::vector<double> lat_edges(int n, const double *wt)
{
    std::vector<double> v(n + 1);
    v[0] = -kPi / 2.0;
    for (int j = 1; j <= n - 1; j++) v[j] = std::asin(wt[j - 1] + std::sin(v[j - 1]));
    v[n] = kPi / 2.0;
    return v;
}

// search_OverwrapRange for one axis, ref :410-429 (first hit of the upper bound ends the scan;
// the lower bound keeps the last hit seen before that).
bool overlap_range(const std::vector<double> &e, int n, double lo, double hi, int &r1, int &r2)
{
    r1 = -1; r2 = -1;
    for (int j = 1; j <= n; j++) {
        if (e[j - 1] <= lo && lo <= e[j]) r1 = j;
        if (e[j - 1] <= hi && hi <= e[j]) { r2 = j; break; }
    }
    return r1 > 0 && r2 > 0;
}

struct LatRow {             // per destination latitude row
    int ry1 = 0, nyr = 0;
    std::vector<double> w1, w2;
};

}  // namespace

extern "C" int dccm_grid_gauss(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt)
{
    if (im < 1 || jm < 1) return fail(DCCM_ERR_ARG, "dccm_grid_gauss: bad size %d x %d", im, jm);
    std::vector<double> mu, w;
    gauss_legendre(jm, mu, w);
    for (int j = 0; j < jm; j++) { y_Lat[j] = std::asin(mu[j]); y_LatWt[j] = w[j]; }
    for (int i = 0; i < im; i++) { x_Lon[i] = 2.0 * kPi * i / im; x_LonWt[i] = 2.0 * kPi / im; }
    return DCCM_OK;
}

extern "C" int dccm_grid_regular(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt)
{
    if (im < 1 || jm < 1) return fail(DCCM_ERR_ARG, "dccm_grid_regular: bad size %d x %d", im, jm);
    for (int j = 0; j < jm; j++) {
        double e0 = -0.5 * kPi + kPi * j / jm, e1 = -0.5 * kPi + kPi * (j + 1) / jm;
        if (j == jm - 1) e1 = 0.5 * kPi;
        y_Lat[j] = 0.5 * (e0 + e1);
        y_LatWt[j] = std::sin(e1) - std::sin(e0);
    }
    for (int i = 0; i < im; i++) { x_Lon[i] = 2.0 * kPi * i / im; x_LonWt[i] = 2.0 * kPi / im; }
    return DCCM_OK;
}

// ref tool/gmapgen/gmapgen_main.f90:336-405.  The O(n^2) exchange sort (:407-426) is replaced
// by std::sort: same ascending result.
extern "C" int dccm_grid_exchange(int jma, const double *y_LatA, const double *y_IntWtLatA,
                                  int jmo, const double *y_IntWtLatO,
                                  int *jms, double *y_LatS, double *y_IntWtLatS)
{
    if (jma < 1 || jmo < 1) return fail(DCCM_ERR_ARG, "dccm_grid_exchange: bad sizes");
    if (jma == jmo) {
        *jms = jmo;
        for (int j = 0; j < jmo; j++) { y_LatS[j] = y_LatA[j]; y_IntWtLatS[j] = y_IntWtLatA[j]; }
        return DCCM_OK;
    }
    std::vector<double> fa = lat_edges(jma, y_IntWtLatA), fo = lat_edges(jmo, y_IntWtLatO);
    std::vector<double> fs;
    fs.reserve(jma + jmo);
    for (int j = 0; j <= jma - 1; j++) fs.push_back(fa[j]);
    for (int j = 1; j <= jmo; j++) fs.push_back(fo[j]);
    std::sort(fs.begin(), fs.end());
    int n = 0;
    for (size_t j = 1; j < fs.size(); j++) {
        double intWt = std::sin(fs[j]) - std::sin(fs[j - 1]);
        if (std::fabs(intWt) > 1e-12) {
            y_LatS[n] = 0.5 * (fs[j - 1] + fs[j]);
            y_IntWtLatS[n] = intWt;
            n++;
        }
    }
    *jms = n;
    return DCCM_OK;
}

extern "C" int dccm_table_gen_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                      int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                      const double *y_LatIntWtS, const double *y_LatIntWtD,
                                      int accuracy_order, int lon_mode, dccm_table **out)
{
    return dccm_table_gen_jones99_rows(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                                       accuracy_order, lon_mode, 1, nyd, out);
}

// sep != nullptr: return the separable factors instead of the expanded table (sep->ok stays false when the
// pair is not handled in that form: equal longitudes / nx == 1 -- those tables are zonal stencils -- or 2nd order)
static int jones99_impl(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                        int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                        const double *y_LatIntWtS, const double *y_LatIntWtD,
                        int accuracy_order, int lon_mode, int jd_first, int jd_last,
                        dccm_table **out, SepFactors *sep)
{
    (void)y_LatD;
    if (out) *out = nullptr;
    if (nxs < 1 || nys < 1 || nxd < 1 || nyd < 1) return fail(DCCM_ERR_ARG, "jones99: bad grid sizes");
    if (jd_first < 1 || jd_last > nyd || jd_first > jd_last + 1) return fail(DCCM_ERR_ARG, "jones99: bad destination row range");
    std::vector<double> uS = lon_edges(nxs, x_LonS), uD = lon_edges(nxd, x_LonD);
    std::vector<double> vS = lat_edges(nys, y_LatIntWtS), vD = lat_edges(nyd, y_LatIntWtD);

    // ---- longitude: per destination column (ref :402-419) ----
    bool general = false;
    std::vector<int> lx1(nxd + 1), lxn(nxd + 1);          // reference mode: first source column, count
    std::vector<int> gptr, gidx; std::vector<double> gw;  // general mode: CSR of (source column, fraction)
    if (nxs == 1 && nxd != 1) {
        for (int i = 1; i <= nxd; i++) { lx1[i] = 1; lxn[i] = 1; }
    } else if (nxs != 1 && nxd == 1) {
        lx1[1] = 1; lxn[1] = nxs;
    } else {
        bool same = (nxs == nxd);
        if (same) for (int i = 0; i < nxs; i++) if (x_LonS[i] != x_LonD[i]) { same = false; break; }
        if (lon_mode == 1 && !same) {
            if (accuracy_order > 1)
                return fail(DCCM_ERR_UNSUPPORTED, "jones99: 2nd order needs equal longitudes or nx==1");
            general = true;
            gptr.assign(nxd + 1, 0);
            for (int id = 1; id <= nxd; id++) {
                double a = uD[id - 1], b = uD[id];
                gptr[id - 1] = (int)gidx.size();
                for (int shift = -1; shift <= 1; shift++) {
                    double off = 2.0 * kPi * shift;
                    // source columns whose shifted cell can touch [a,b]
                    for (int m = 1; m <= nxs; m++) {
                        double lo = uS[m - 1] + off, hi = uS[m] + off;
                        double l = lo > a ? lo : a, h = hi < b ? hi : b;
                        double ov = h - l;
                        if (ov > 0.0) { gidx.push_back(m); gw.push_back(ov / (b - a)); }
                    }
                }
            }
            gptr[nxd] = (int)gidx.size();
        } else {
            for (int id = 1; id <= nxd; id++) {
                int r1, r2;
                if (!overlap_range(uS, nxs, uD[id - 1], uD[id], r1, r2))
                    return fail(DCCM_ERR_UNSUPPORTED,
                                "jones99: longitude overlap search failed at iD=%d; the reference generator "
                                "needs equal longitudes or nx==1 (use lon_mode=1)", id);
                lx1[id] = r1; lxn[id] = r2 - r1 + 1;
            }
        }
    }

    // ---- latitude: per destination row (ref :421-438, :314-365) ----
    const double DLon_k = 2.0 * kPi / (double)nxd;
    double DLon_nk, DLon_n;
    if (nxd == 1) { DLon_nk = 2.0 * kPi / (double)nxs; DLon_n = DLon_nk; }
    else          { DLon_nk = 2.0 * kPi / (double)nxd; DLon_n = 2.0 * kPi; }
    std::vector<LatRow> rows(nyd + 1);
    std::vector<double> seg(nys + 2);
    for (int jD = jd_first; jD <= jd_last; jD++) {
        int r1, r2;
        if (!overlap_range(vS, nys, vD[jD - 1], vD[jD], r1, r2))
            return fail(DCCM_ERR_SEARCH, "jones99: latitude overlap search failed at jD=%d", jD);
        LatRow &R = rows[jD];
        R.ry1 = r1; R.nyr = r2 - r1 + 1;
        R.w1.assign(R.nyr + 1, 0.0);
        seg[0] = vD[jD - 1];
        for (int j = 1; j <= R.nyr - 1; j++) seg[j] = vS[r1 + j - 1];
        seg[R.nyr] = vD[jD];
        double lat1_k = vD[jD - 1], lat2_k = vD[jD];
        double Ak = (std::sin(lat2_k) - std::sin(lat1_k)) * DLon_k;
        for (int j = 1; j <= R.nyr; j++) {
            double a = seg[j - 1], b = seg[j];
            if (general) R.w1[j] = (std::sin(b) - std::sin(a)) / (std::sin(lat2_k) - std::sin(lat1_k));
            else         R.w1[j] = DLon_nk * (std::sin(b) - std::sin(a)) / Ak;
        }
        if (accuracy_order > 1) {
            R.w2.assign(R.nyr + 1, 0.0);
            for (int j = 1; j <= R.nyr; j++) {
                double a = seg[j - 1], b = seg[j];
                double na = vS[r1 + j - 2], nb = vS[r1 + j - 1];
                double An = (std::sin(nb) - std::sin(na)) * DLon_n;
                R.w2[j] = ((std::cos(b) + b * std::sin(b)) - (std::cos(a) + a * std::sin(a))) * DLon_nk / Ak
                        - ((std::cos(nb) + nb * std::sin(nb)) - (std::cos(na) + na * std::sin(na))) * DLon_n * R.w1[j] / An;
            }
        }
    }

    if (sep && !general) {
        // every destination column must see the same longitude shift(s): true for equal longitudes and nx == 1
        sep->nxs = nxs; sep->nys = nys; sep->nxd = nxd; sep->nyd = nyd;
        for (int iD = 2; iD <= nxd; iD++)
            if (lxn[iD] != lxn[1] || (nxs > 1 && (lx1[iD] - iD - (lx1[1] - 1)) % nxs != 0)) return DCCM_OK;
        sep->zptr.assign(nyd + 1, 0);
        for (int jD = 1; jD <= nyd; jD++) {
            const LatRow &R = rows[jD];
            for (int m = 1; m <= lxn[1]; m++) {
                const int di = (lx1[1] + m - 2) % nxs;                  // column 1 (0-based 0): shift = source column
                for (int n = 1; n <= R.nyr; n++) {
                    const int jS = R.ry1 + n - 1;
                    if (std::fabs(R.w1[n]) > 1e-14) { sep->zdi.push_back(di); sep->zjs.push_back(jS - 1); sep->zw.push_back(R.w1[n]); }
                    if (accuracy_order > 1) {
                        int j1, j2;
                        if (jS == 1)        { j1 = jS;     j2 = jS + 1; }
                        else if (jS == nys) { j1 = jS - 1; j2 = jS; }
                        else                { j1 = jS - 1; j2 = jS + 1; }
                        const double DLat = y_LatS[j2 - 1] - y_LatS[j1 - 1];
                        if (std::fabs(R.w2[n]) > 1e-14) {
                            sep->zdi.push_back(di); sep->zjs.push_back(j1 - 1); sep->zw.push_back(-R.w2[n] / DLat);
                            sep->zdi.push_back(di); sep->zjs.push_back(j2 - 1); sep->zw.push_back(+R.w2[n] / DLat);
                        }
                    }
                }
            }
            sep->zptr[jD] = (int32_t)sep->zw.size();
        }
        sep->zonal = true;
        return DCCM_OK;
    }
    if (sep) {
        if (accuracy_order > 1) return DCCM_OK;
        sep->mode = 0; sep->nxs = nxs; sep->nys = nys; sep->nxd = nxd; sep->nyd = nyd;
        sep->xptr.assign(gptr.begin(), gptr.end());
        sep->xi.resize(gidx.size());
        for (size_t k = 0; k < gidx.size(); k++) sep->xi[k] = gidx[k] - 1;
        sep->xw = gw;
        sep->yptr.assign(nyd + 1, 0);
        for (int jD = 1; jD <= nyd; jD++) {
            const LatRow &R = rows[jD];
            for (int n = 1; n <= R.nyr; n++) { sep->yj.push_back(R.ry1 + n - 2); sep->yw.push_back(R.w1[n]); }
            sep->yptr[jD] = (int32_t)sep->yj.size();
        }
        sep->ok = true;
        return DCCM_OK;
    }

    // ---- emit in table-file order (ref :230-275) ----
    dccm_table *t = new dccm_table();
    for (int jD = jd_first; jD <= jd_last; jD++) {
        const LatRow &R = rows[jD];
        for (int iD = 1; iD <= nxd; iD++) {
            int nxr = general ? gptr[iD] - gptr[iD - 1] : lxn[iD];
            for (int m = 1; m <= nxr; m++) {
                int iS = general ? gidx[gptr[iD - 1] + m - 1] : lx1[iD] + m - 1;
                for (int n = 1; n <= R.nyr; n++) {
                    int jS = R.ry1 + n - 1;
                    double w = general ? gw[gptr[iD - 1] + m - 1] * R.w1[n] : R.w1[n];
                    if (std::fabs(w) > 1e-14) t->push(iD, jD, iS, jS, w);
                    if (accuracy_order > 1) {
                        int j1, j2;
                        if (jS == 1)        { j1 = jS;     j2 = jS + 1; }
                        else if (jS == nys) { j1 = jS - 1; j2 = jS; }
                        else                { j1 = jS - 1; j2 = jS + 1; }
                        double DLat = y_LatS[j2 - 1] - y_LatS[j1 - 1];
                        if (std::fabs(R.w2[n]) > 1e-14) {
                            t->push(iD, jD, iS, j1, -R.w2[n] / DLat);
                            t->push(iD, jD, iS, j2, +R.w2[n] / DLat);
                        }
                    }
                }
            }
        }
    }
    *out = t;
    return DCCM_OK;
}

extern "C" int dccm_table_gen_jones99_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                           int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                           const double *y_LatIntWtS, const double *y_LatIntWtD,
                                           int accuracy_order, int lon_mode, int jd_first, int jd_last,
                                           dccm_table **out)
{
    return jones99_impl(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                        accuracy_order, lon_mode, jd_first, jd_last, out, nullptr);
}

int dccm::jones99_factors(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                          int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                          const double *y_LatIntWtS, const double *y_LatIntWtD,
                          int accuracy_order, int lon_mode, SepFactors &f)
{
    f = SepFactors();
    return jones99_impl(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                        accuracy_order, lon_mode, 1, nyd, nullptr, &f);
}

extern "C" int dccm_table_gen_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                       int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                       int lon_mode, dccm_table **out)
{
    return dccm_table_gen_bilinear_rows(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, 1, nyr, out);
}

static int bilinear_impl(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                         int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                         int lon_mode, int jr_first, int jr_last, dccm_table **out, SepFactors *sep)
{
    if (out) *out = nullptr;
    if (nxs < 1 || nys < 2 || nxr < 1 || nyr < 1) return fail(DCCM_ERR_ARG, "bilinear: bad grid sizes");
    if (jr_first < 1 || jr_last > nyr || jr_first > jr_last + 1) return fail(DCCM_ERR_ARG, "bilinear: bad destination row range");
    const double dlon_r = 360.0 / (double)nxr, dlon_s = 360.0 / (double)nxs;   // ref :78-79
    dccm_table *t = new dccm_table();
    if (!sep) t->reserve((size_t)nxr * (jr_last - jr_first + 1) * ((nxr == 1) ? 2 * nxs : (nxs == 1 ? 2 : 4)));
    // longitude part is the same for every row: hoist (ref :116-119, cal_coef :154-165)
    std::vector<int> is1(nxr), is2(nxr);
    std::vector<double> a1(nxr), a2(nxr);
    if (nxr != 1 && nxs != 1) {
        for (int ir = 1; ir <= nxr; ir++) {
            int is = (int)(dlon_r * (ir - 1) / dlon_s) + 1;
            int ie = is % nxs + 1;
            double x1 = x_LonS[is - 1], x3 = x_LonS[ie - 1], xc = x_LonR[ir - 1];
            if (lon_mode == 1 && x3 <= x1) x3 += 2.0 * kPi;
            is1[ir - 1] = is; is2[ir - 1] = ie;
            a1[ir - 1] = (xc - x1) / (x3 - x1);
            a2[ir - 1] = 1.0 - a1[ir - 1];
        }
    }
    bool zon = false;
    if (sep && nxr != 1 && nxs != 1 && nxr == nxs) {
        zon = true;                                   // same shifts and value-equal coefficients in every column?
        for (int ir = 1; ir < nxr && zon; ir++)
            zon = (is1[ir] - 1 - ir + nxs) % nxs == (is1[0] - 1) % nxs && (is2[ir] - 1 - ir + nxs) % nxs == (is2[0] - 1) % nxs &&
                  a1[ir] == a1[0] && a2[ir] == a2[0];
    }
    if (sep) {
        if (nxr == 1 || nxs == 1) { delete t; return DCCM_OK; }     // axisymmetric side: goes through the table
        sep->mode = 1; sep->nxs = nxs; sep->nys = nys; sep->nxd = nxr; sep->nyd = nyr;
        sep->xptr.resize(nxr + 1); sep->yptr.resize(nyr + 1);
        for (int ir = 0; ir < nxr; ir++) {
            sep->xptr[ir] = 2 * ir;
            sep->xi.push_back(is1[ir] - 1); sep->xw.push_back(a2[ir]);       // m0: (is,  a2)
            sep->xi.push_back(is2[ir] - 1); sep->xw.push_back(a1[ir]);       // m1: (is+1, a1)
        }
        sep->xptr[nxr] = 2 * nxr;
    }
    for (int jr = jr_first; jr <= jr_last; jr++) {
        double latR = y_LatR[jr - 1];
        int js = -1; bool extp = true;                                   // get_correspondID_latS :130-152
        for (int j = 1; j <= nys - 1; j++)
            if (y_LatS[j - 1] < latR && latR <= y_LatS[j]) { js = j; extp = false; break; }
        if (extp) {
            if (latR <= y_LatS[0]) js = 1;
            if (latR > y_LatS[nys - 1]) js = nys;
            if (js < 0) { delete t; return fail(DCCM_ERR_SEARCH, "bilinear: NaN latitude at row %d", jr); }
        }
        if (nxr == 1) {
            if (!extp) {
                double b1 = (latR - y_LatS[js - 1]) / (y_LatS[js] - y_LatS[js - 1]);
                double c1 = (1.0 - b1) / nxs, c2 = b1 / nxs;
                for (int is = 1; is <= nxs; is++) { t->push(1, jr, is, js, c1); t->push(1, jr, is, js % nys + 1, c2); }
            } else {
                for (int is = 1; is <= nxs; is++) t->push(1, jr, is, js, 1.0 / nxs);
            }
        } else if (nxs == 1) {
            if (!extp) {
                double b1 = (latR - y_LatS[js - 1]) / (y_LatS[js] - y_LatS[js - 1]);
                double c1 = 1.0 - b1, c2 = b1;
                for (int ir = 1; ir <= nxr; ir++) { t->push(ir, jr, 1, js, c1); t->push(ir, jr, 1, js % nys + 1, c2); }
            } else {
                for (int ir = 1; ir <= nxr; ir++) t->push(ir, jr, 1, js, 1.0);
            }
        } else {
            // The reference does not test extp_flag here; js = nys would read y_LatS(nys+1).
            // Mirror the southern edge: extrapolate from the last two rows.
            int jlo = js, jhi = js + 1;
            if (js == nys) { jlo = nys - 1; jhi = nys; }
            int jhi_out = (js == nys) ? jhi : (js % nys + 1);
            double y1 = y_LatS[jlo - 1], y3 = y_LatS[jhi - 1];
            double b1 = (latR - y1) / (y3 - y1), b2 = 1.0 - b1;
            if (sep) {
                sep->yptr[jr - 1] = 2 * (jr - 1);
                sep->yj.push_back(jlo - 1);     sep->yw.push_back(b2);      // n0: (jlo, b2)
                sep->yj.push_back(jhi_out - 1); sep->yw.push_back(b1);      // n1: (jhi, b1)
                if (zon) {                                                  // the row's stencil, in table order
                    const int d1 = (is1[0] - 1) % nxs, d2 = (is2[0] - 1) % nxs;
                    const int dd[4] = {d1, d2, d2, d1}, jj[4] = {jlo - 1, jlo - 1, jhi_out - 1, jhi_out - 1};
                    const double ww[4] = {a2[0] * b2, a1[0] * b2, a1[0] * b1, a2[0] * b1};
                    for (int k = 0; k < 4; k++) { sep->zdi.push_back(dd[k]); sep->zjs.push_back(jj[k]); sep->zw.push_back(ww[k]); }
                }
                continue;
            }
            for (int ir = 0; ir < nxr; ir++) {
                t->push(ir + 1, jr, is1[ir], jlo,     a2[ir] * b2);
                t->push(ir + 1, jr, is2[ir], jlo,     a1[ir] * b2);
                t->push(ir + 1, jr, is2[ir], jhi_out, a1[ir] * b1);
                t->push(ir + 1, jr, is1[ir], jhi_out, a2[ir] * b1);
            }
        }
    }
    if (sep) {
        sep->yptr[nyr] = 2 * nyr;
        sep->ok = true;
        if (zon) {
            sep->zptr.resize(nyr + 1);
            for (int jr = 0; jr <= nyr; jr++) sep->zptr[jr] = 4 * jr;
            sep->zonal = true;
        }
        delete t;
        return DCCM_OK;
    }
    *out = t;
    return DCCM_OK;
}

extern "C" int dccm_table_gen_bilinear_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                            int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                            int lon_mode, int jr_first, int jr_last, dccm_table **out)
{
    return bilinear_impl(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, jr_first, jr_last, out, nullptr);
}

// make_mapping_table, ref common/cal_mappingtable.f90:10-49 (+ cal_coef :59-76): the stand-alone bilinear generator on
// regular grids given by their sizes only (degrees; longitudes from 0 in steps of 360/nx, latitudes from one pole to
// the other in steps of 180/(ny-1)).  The west source column and its weight pair depend only on the destination
// column, the south source row and its pair only on the destination row, so both are tabulated once (the reference
// recomputes them per cell); a cell then emits the four products in the reference order, keeping those > 0.
extern "C" int dccm_table_gen_make_mapping_table(int nx_r, int ny_r, int nx_s, int ny_s, dccm_table **out)
{
    if (!out) return fail(DCCM_ERR_ARG, "make_mapping_table: out is NULL");
    *out = nullptr;
    if (nx_r < 1 || nx_s < 1 || ny_r < 2 || ny_s < 2)
        return fail(DCCM_ERR_ARG, "make_mapping_table: need nx >= 1 and ny >= 2 on both sides (got %d x %d <- %d x %d)",
                    nx_r, ny_r, nx_s, ny_s);
    struct Axis { int lo, hi; double w_hi, w_lo; };          // source index pair (1-based) and weights (a1 | a2)
    auto tabulate = [](int n_r, double d_r, int n_s, double d_s) {
        std::vector<Axis> ax(n_r);
        for (int k = 0; k < n_r; k++) {
            const double c = d_r * k;
            const int lo = (int)(c / d_s) + 1;
            const double c1 = (lo - 1) * d_s, c3 = c1 + d_s;
            const double t = (c - c1) / (c3 - c1);
            ax[k] = {lo, lo % n_s + 1, t, 1.0 - t};
        }
        return ax;
    };
    const std::vector<Axis> X = tabulate(nx_r, 360.0 / nx_r, nx_s, 360.0 / nx_s);
    const std::vector<Axis> Y = tabulate(ny_r, 180.0 / (ny_r - 1), ny_s, 180.0 / (ny_s - 1));
    auto *t = new dccm_table;
    t->reserve((size_t)4 * nx_r * ny_r);
    for (int j = 0; j < ny_r; j++) {
        const Axis &y = Y[j];
        for (int i = 0; i < nx_r; i++) {
            const Axis &x = X[i];
            const double sw = x.w_lo * y.w_lo, se = x.w_hi * y.w_lo, ne = x.w_hi * y.w_hi, nw = x.w_lo * y.w_hi;
            if (sw > 0.0) t->push(i + 1, j + 1, x.lo, y.lo, sw);
            if (se > 0.0) t->push(i + 1, j + 1, x.hi, y.lo, se);
            if (ne > 0.0) t->push(i + 1, j + 1, x.hi, y.hi, ne);
            if (nw > 0.0) t->push(i + 1, j + 1, x.lo, y.hi, nw);
        }
    }
    *out = t;
    return DCCM_OK;
}

int dccm::bilinear_factors(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                           int nxr, const double *x_LonR, int nyr, const double *y_LatR, int lon_mode, SepFactors &f)
{
    f = SepFactors();
    return bilinear_impl(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, 1, nyr, nullptr, &f);
}

int dccm::slice_rows(SepFactors &f, int j0, int j1, int src_row0, int src_rows)
{
    auto cut = [&](std::vector<int32_t> &ptr, std::vector<int32_t> &rows, std::vector<double> &w,
                   std::vector<int32_t> *extra) -> int {
        if (ptr.empty()) return DCCM_OK;
        const int e0 = ptr[j0], e1 = ptr[j1];
        for (int e = e0; e < e1; e++) {
            // padding entries of weight 0 may point anywhere; real entries must lie inside the band's source buffer
            const int r = rows[e] - src_row0;
            if ((r < 0 || r >= src_rows) && w[e] != 0.0)
                return fail(DCCM_ERR_ARG, "dccm_remap_create (band): destination rows [%d,%d) read source row %d outside [%d,%d)",
                            j0, j1, rows[e], src_row0, src_row0 + src_rows);
            rows[e] = (r < 0 || r >= src_rows) ? 0 : r;
        }
        std::vector<int32_t> np(ptr.begin() + j0, ptr.begin() + j1 + 1);
        for (auto &v : np) v -= e0;
        ptr.swap(np);
        rows.assign(rows.begin() + e0, rows.begin() + e1);
        w.assign(w.begin() + e0, w.begin() + e1);
        if (extra) extra->assign(extra->begin() + e0, extra->begin() + e1);
        return DCCM_OK;
    };
    int rc = cut(f.yptr, f.yj, f.yw, nullptr);
    if (rc) return rc;
    if (f.zonal) { rc = cut(f.zptr, f.zjs, f.zw, &f.zdi); if (rc) return rc; }
    f.nyd = j1 - j0; f.nys = src_rows;
    return DCCM_OK;
}


// Host-side expansion of the separable factors into a table, entry by entry as the kernels do it (kind 2):
// lets the CPU test-suite check factors + expansion order against the generators index for index.
static int expand_zonal(const SepFactors &f, dccm_table **out)
{
    dccm_table *t = new dccm_table();
    for (int jD = 0; jD < f.nyd; jD++)
        for (int iD = 0; iD < f.nxd; iD++)
            for (int e = f.zptr[jD]; e < f.zptr[jD + 1]; e++)
                t->push(iD + 1, jD + 1, (f.nxs == 1 ? 0 : (iD + f.zdi[e]) % f.nxs) + 1, f.zjs[e] + 1, f.zw[e]);
    *out = t;
    return DCCM_OK;
}

static int expand_factors(const SepFactors &f, dccm_table **out)
{
    if (f.zonal) return expand_zonal(f, out);
    if (!f.ok) return fail(DCCM_ERR_UNSUPPORTED, "this grid pair is not handled in separable form (equal longitudes, nx == 1 or 2nd order)");
    dccm_table *t = new dccm_table();
    for (int jD = 0; jD < f.nyd; jD++) {
        const int y0 = f.yptr[jD], y1 = f.yptr[jD + 1];
        for (int iD = 0; iD < f.nxd; iD++) {
            const int x0 = f.xptr[iD], x1 = f.xptr[iD + 1];
            if (f.mode == 1) {
                const int mm[4] = {x0, x0 + 1, x0 + 1, x0}, nn[4] = {y0, y0, y0 + 1, y0 + 1};
                for (int k = 0; k < 4; k++)
                    t->push(iD + 1, jD + 1, f.xi[mm[k]] + 1, f.yj[nn[k]] + 1, f.xw[mm[k]] * f.yw[nn[k]]);
            } else {
                for (int m = x0; m < x1; m++)
                    for (int n = y0; n < y1; n++) {
                        const double w = f.xw[m] * f.yw[n];
                        if (std::fabs(w) > 1e-14) t->push(iD + 1, jD + 1, f.xi[m] + 1, f.yj[n] + 1, w);
                    }
            }
        }
    }
    *out = t;
    return DCCM_OK;
}

extern "C" int dccm_table_gen_jones99_separable(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                                int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                                const double *y_LatIntWtS, const double *y_LatIntWtD,
                                                int accuracy_order, int lon_mode, dccm_table **out)
{
    *out = nullptr;
    SepFactors f;
    int rc = jones99_factors(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                             accuracy_order, lon_mode, f);
    return rc ? rc : expand_factors(f, out);
}

extern "C" int dccm_table_gen_bilinear_separable(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                                 int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                                 int lon_mode, dccm_table **out)
{
    *out = nullptr;
    SepFactors f;
    int rc = bilinear_factors(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, f);
    return rc ? rc : expand_factors(f, out);
}

// host-only check of the band operators (dccm_remap_create_*_band): the band's factors multiplied out, LOCAL indices
extern "C" int dccm_table_gen_band_expanded(int conservative, int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                            int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                            const double *y_LatIntWtS, const double *y_LatIntWtD,
                                            int accuracy_order, int lon_mode,
                                            int jD0, int jD1, int src_row0, int src_rows, dccm_table **out)
{
    *out = nullptr;
    SepFactors f;
    int rc = conservative ? jones99_factors(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                                            accuracy_order, lon_mode, f)
                          : bilinear_factors(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, lon_mode, f);
    if (rc) return rc;
    if (jD0 < 0 || jD1 > nyd || jD0 >= jD1) return fail(DCCM_ERR_ARG, "dccm_table_gen_band_expanded: bad rows");
    rc = slice_rows(f, jD0, jD1, src_row0, src_rows);
    return rc ? rc : expand_factors(f, out);
}

extern "C" int dccm_table_write_text(const dccm_table *t, const char *filename)
{
    FILE *f = fopen(filename, "w");
    if (!f) return fail(DCCM_ERR_IO, "cannot open %s for writing", filename);
    std::vector<char> buf(1 << 20);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (size_t k = 0; k < t->coef.size(); k++)
        fprintf(f, "%12d%12d%12d%12d  %24.16E\n", t->iD[k], t->jD[k], t->iS[k], t->jS[k], t->coef[k]);
    fclose(f);
    return DCCM_OK;
}

// list-directed read of "ir jr is js coef" (ref common/grid_mapping_util_jones99.f90:479-504);
// blanks or commas separate, D exponents accepted; blank lines are skipped as a list-directed read skips them, and so
// is a line that does not parse (the reference would die of a run-time I/O error there); end of file ends the read
// like the reference's end=200.
extern "C" int dccm_table_read_text(const char *filename, dccm_table **out)
{
    *out = nullptr;
    FILE *f = fopen(filename, "r");
    if (!f) return fail(DCCM_ERR_IO, "cannot open %s", filename);
    dccm_table *t = new dccm_table();
    char line[512];
    while (fgets(line, sizeof line, f)) {
        for (char *p = line; *p; p++) {
            if (*p == ',') *p = ' ';
            if (*p == 'D' || *p == 'd') *p = 'E';
        }
        char *p = line, *e;
        long v[4];
        bool ok = true;
        for (int k = 0; k < 4; k++) {
            v[k] = strtol(p, &e, 10);
            if (e == p) { ok = false; break; }
            p = e;
        }
        if (!ok) continue;
        double c = strtod(p, &e);
        if (e == p) continue;
        t->push((int)v[0], (int)v[1], (int)v[2], (int)v[3], c);
    }
    fclose(f);
    *out = t;
    return DCCM_OK;
}

static const char kMagic[8] = {'D', 'C', 'C', 'M', 'T', 'B', 'L', '1'};

extern "C" int dccm_table_write_bin(const dccm_table *t, const char *filename)
{
    FILE *f = fopen(filename, "wb");
    if (!f) return fail(DCCM_ERR_IO, "cannot open %s for writing", filename);
    int64_t n = (int64_t)t->coef.size();
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1;
    ok = ok && fwrite(t->iD.data(), 4, n, f) == (size_t)n && fwrite(t->jD.data(), 4, n, f) == (size_t)n;
    ok = ok && fwrite(t->iS.data(), 4, n, f) == (size_t)n && fwrite(t->jS.data(), 4, n, f) == (size_t)n;
    ok = ok && fwrite(t->coef.data(), 8, n, f) == (size_t)n;
    fclose(f);
    return ok ? DCCM_OK : fail(DCCM_ERR_IO, "short write to %s", filename);
}

extern "C" int dccm_table_read_bin(const char *filename, dccm_table **out)
{
    *out = nullptr;
    FILE *f = fopen(filename, "rb");
    if (!f) return fail(DCCM_ERR_IO, "cannot open %s", filename);
    char magic[8];
    int64_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(&n, 8, 1, f) != 1 || n < 0) {
        fclose(f);
        return fail(DCCM_ERR_IO, "%s is not a dccm binary table", filename);
    }
    dccm_table *t = new dccm_table();
    t->iD.resize(n); t->jD.resize(n); t->iS.resize(n); t->jS.resize(n); t->coef.resize(n);
    bool ok = fread(t->iD.data(), 4, n, f) == (size_t)n && fread(t->jD.data(), 4, n, f) == (size_t)n
           && fread(t->iS.data(), 4, n, f) == (size_t)n && fread(t->jS.data(), 4, n, f) == (size_t)n
           && fread(t->coef.data(), 8, n, f) == (size_t)n;
    fclose(f);
    if (!ok) { delete t; return fail(DCCM_ERR_IO, "short read from %s", filename); }
    *out = t;
    return DCCM_OK;
}

extern "C" int64_t dccm_table_size(const dccm_table *t) { return t ? (int64_t)t->coef.size() : -1; }

extern "C" int dccm_table_get(const dccm_table *t, int32_t *iD, int32_t *jD, int32_t *iS, int32_t *jS, double *coef)
{
    size_t n = t->coef.size();
    if (iD) memcpy(iD, t->iD.data(), 4 * n);
    if (jD) memcpy(jD, t->jD.data(), 4 * n);
    if (iS) memcpy(iS, t->iS.data(), 4 * n);
    if (jS) memcpy(jS, t->jS.data(), 4 * n);
    if (coef) memcpy(coef, t->coef.data(), 8 * n);
    return DCCM_OK;
}

extern "C" int dccm_table_index(const dccm_table *t, int gnxs, int gnxr,
                                int32_t *send_index, int32_t *recv_index, double *coef_s)
{
    size_t n = t->coef.size();
    for (size_t k = 0; k < n; k++) {
        int64_t r = (int64_t)t->iD[k] + (int64_t)gnxr * (t->jD[k] - 1);
        int64_t s = (int64_t)t->iS[k] + (int64_t)gnxs * (t->jS[k] - 1);
        if (r > INT32_MAX || s > INT32_MAX) return fail(DCCM_ERR_ARG, "table index overflows default INTEGER");
        recv_index[k] = (int32_t)r;
        send_index[k] = (int32_t)s;
        coef_s[k] = t->coef[k];
    }
    return DCCM_OK;
}

extern "C" void dccm_table_free(dccm_table *t) { delete t; }
