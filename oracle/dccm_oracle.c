/*
 * dccm_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See dccm_oracle.h.
 *
 * PARITY UNPINNED (no reference tests / golden vectors / buildable reference; see header).
 *
 * Every function restates one reference loop nest; `ref:` gives file:line under
 * /root/reference.  Compile with strict IEEE semantics:
 *     gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math
 * (the reference is built with `-fp-model source`, sysdep/Makedef.Linux64-intel-impi:7-10).
 *
 * Documented deviations from the literal Fortran (all are places where the reference
 * itself is undefined; the product code makes the same choices):
 *   B-1  LatHFlx(:,:,3) after the implicit correction reads a_LatentHeatLocal(3) out of
 *        bounds (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:181,370,379).  We store the
 *        area-weighted composite of the corrected slots 1,2 instead.  Not consumed downstream.
 *   B-3  xya_SfcMOLength (0/0 when CalcFlag is false, :545-550) is a local that is never
 *        used; it is not computed.
 *   C-1  Solve_TriDiagSystem_Forward divides by Mtx(k=1,-1)=0 at k=1
 *        (ref atm/dcpam_sfc_implicit_coupling_mod.f90:393-400 with :208,237,267).  The k=1
 *        row is never read again (Coef1/2 use k=2; RHS(1) is overwritten by the glue,
 *        atm/dccm_atm_mod.f90:832-835), so the sweep stops at k=2: row 1 of the matrices
 *        stays as built and RHS(1) stays the un-swept flux divergence.
 *   A5-1 the bilinear generator's general branch ignores extp_flag and reads y_LatS(nys+1)
 *        out of bounds when a destination latitude lies north of the last source latitude
 *        (ref common/grid_mapping_util.f90:114-118 with :147-151).  We mirror what the same
 *        code does at the southern edge: linear extrapolation from the last two source rows.
 */
#define _GNU_SOURCE
#include "dccm_oracle.h"
#include "orc_pmath.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* exp / log / x**y of the bulk flux and of the atmosphere's fourth root: 1 (default) = the portable
 * definitions of orc_pmath.h (fixed IEEE sequences; the CUDA path evaluates the same sequences, so the two
 * can be compared bit for bit), 0 = the C library's (the reading a Fortran compiler's run-time gives; used
 * to MEASURE how far the portable definitions are from libm over a whole exchange). */
static int g_math = 1;
void orc_set_math(int portable) { g_math = portable ? 1 : 0; }
int orc_get_math(void) { return g_math; }
static inline double m_exp(double x) { return g_math ? orc_pm_exp(x) : exp(x); }
static inline double m_log(double x) { return g_math ? orc_pm_log(x) : log(x); }
static inline double m_pow(double x, double y) { return g_math ? orc_pm_pow(x, y) : pow(x, y); }

void orc_pm_exp_v(int64_t n, const double *x, double *y) { for (int64_t i = 0; i < n; i++) y[i] = orc_pm_exp(x[i]); }
void orc_pm_log_v(int64_t n, const double *x, double *y) { for (int64_t i = 0; i < n; i++) y[i] = orc_pm_log(x[i]); }
void orc_pm_pow_v(int64_t n, const double *x, double e, double *y) { for (int64_t i = 0; i < n; i++) y[i] = orc_pm_pow(x[i], e); }

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* threads used by the `!$omp` regions from now on (the reference's SFC / OCN components run on one rank) */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ table container */

orc_table *orc_table_new(void) { return (orc_table *)calloc(1, sizeof(orc_table)); }

void orc_table_free(orc_table *t)
{
    if (!t) return;
    free(t->iD); free(t->jD); free(t->iS); free(t->jS); free(t->coef);
    free(t);
}

int64_t orc_table_n(const orc_table *t) { return t->n; }

void orc_table_copy(const orc_table *t, int32_t *iD, int32_t *jD, int32_t *iS, int32_t *jS, double *coef)
{
    memcpy(iD, t->iD, sizeof(int32_t) * t->n);
    memcpy(jD, t->jD, sizeof(int32_t) * t->n);
    memcpy(iS, t->iS, sizeof(int32_t) * t->n);
    memcpy(jS, t->jS, sizeof(int32_t) * t->n);
    memcpy(coef, t->coef, sizeof(double) * t->n);
}

static void table_push(orc_table *t, int iD, int jD, int iS, int jS, double c)
{
    if (t->n == t->cap) {
        int64_t cap = t->cap ? t->cap * 2 : 4096;
        t->iD = (int32_t *)realloc(t->iD, sizeof(int32_t) * cap);
        t->jD = (int32_t *)realloc(t->jD, sizeof(int32_t) * cap);
        t->iS = (int32_t *)realloc(t->iS, sizeof(int32_t) * cap);
        t->jS = (int32_t *)realloc(t->jS, sizeof(int32_t) * cap);
        t->coef = (double *)realloc(t->coef, sizeof(double) * cap);
        t->cap = cap;
    }
    t->iD[t->n] = iD; t->jD[t->n] = jD; t->iS[t->n] = iS; t->jS[t->n] = jS; t->coef[t->n] = c;
    t->n++;
}

/* ------------------------------------------------------------------ grids */

/* Gauss-Legendre nodes (ascending, south to north) and weights (sum = 2).
 * Stand-in for SPML w_module's xy_Lat / y_Lat_Weight (ref tool/gmapgen/gmapgen_main.f90:256-307);
 * SPML is not vendored, so this supplier is synthesised. */
int orc_gauss_legendre(int n, double *mu, double *w)
{
    const double PI = acos(-1.0);
    for (int i = 0; i < (n + 1) / 2; i++) {
        double x = cos(PI * (i + 0.75) / (n + 0.5));
        double dp = 1.0;
        for (int it = 0; it < 100; it++) {
            double p0 = 1.0, p1 = x;
            for (int l = 2; l <= n; l++) {
                double p2 = ((2.0 * l - 1.0) * x * p1 - (l - 1.0) * p0) / l;
                p0 = p1; p1 = p2;
            }
            if (n == 1) { p0 = 1.0; p1 = x; }
            dp = n * (x * p1 - p0) / (x * x - 1.0);
            double dx = p1 / dp;
            x -= dx;
            if (fabs(dx) < 1e-16) break;
        }
        /* recompute derivative at the converged node */
        {
            double p0 = 1.0, p1 = x;
            for (int l = 2; l <= n; l++) {
                double p2 = ((2.0 * l - 1.0) * x * p1 - (l - 1.0) * p0) / l;
                p0 = p1; p1 = p2;
            }
            dp = n * (x * p1 - p0) / (x * x - 1.0);
        }
        double wi = 2.0 / ((1.0 - x * x) * dp * dp);
        mu[i] = -x; mu[n - 1 - i] = x;
        w[i] = wi;  w[n - 1 - i] = wi;
    }
    if (n % 2 == 1) mu[n / 2] = 0.0;
    return 0;
}

/* Gaussian grid in the form gmapgen hands to the generators (radians, S->N, lon from 0). */
int orc_gauss_grid(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt)
{
    const double PI = acos(-1.0);
    double *mu = (double *)malloc(sizeof(double) * jm);
    orc_gauss_legendre(jm, mu, y_LatWt);
    for (int j = 0; j < jm; j++) y_Lat[j] = asin(mu[j]);
    for (int i = 0; i < im; i++) { x_Lon[i] = 2.0 * PI * i / im; x_LonWt[i] = 2.0 * PI / im; }
    free(mu);
    return 0;
}

/* Regular lat-lon grid (cell centres; weight = sin(north edge) - sin(south edge), sum = 2). */
int orc_regular_grid(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt)
{
    const double PI = acos(-1.0);
    for (int j = 0; j < jm; j++) {
        double e0 = -0.5 * PI + PI * j / jm, e1 = -0.5 * PI + PI * (j + 1) / jm;
        if (j == jm - 1) e1 = 0.5 * PI;
        y_Lat[j] = 0.5 * (e0 + e1);
        y_LatWt[j] = sin(e1) - sin(e0);
    }
    for (int i = 0; i < im; i++) { x_Lon[i] = 2.0 * PI * i / im; x_LonWt[i] = 2.0 * PI / im; }
    return 0;
}

/* ------------------------------------------------------------------ Jones (1999) generator */

/* ref common/grid_mapping_util_jones99.f90:116-152 (calc_edge_coordinate).
 * x_C has n+2 entries (halo), x_F gets n+1 edges x_F[0..n]; y_F gets n+1 edges. */
static void calc_lon_edges(int n, const double *x_Lon, double *u)
{
    const double PI = acos(-1.0);
    double *xc = (double *)malloc(sizeof(double) * (n + 2));
    for (int i = 0; i < n; i++) xc[i + 1] = x_Lon[i];
    xc[0] = xc[n] - 2.0 * PI;          /* :91 */
    xc[n + 1] = xc[1] + 2.0 * PI;      /* :92 */
    for (int i = 0; i <= n; i++) u[i] = 0.5 * (xc[i] + xc[i + 1]);   /* :142-144 */
    free(xc);
}

static void calc_lat_edges(int n, const double *wt, double *v)
{
    const double PI = acos(-1.0);
    v[0] = -PI / 2.0;                                               /* :147 */
    for (int j = 1; j <= n - 1; j++) v[j] = asin(wt[j - 1] + sin(v[j - 1]));   /* :148-150 */
    v[n] = PI / 2.0;                                                /* :151 */
}

/* ref common/grid_mapping_util_jones99.f90:156-442 (gen_gridmapfile_lonlat2lonlatCore),
 * with the driver part :35-114.  Entries are appended in table-file order (:230-275). */
/* jD0..jD1 (1-based, inclusive): the destination rows to emit -- the reference always emits 1..nyd; a sub-range gives
 * exactly the lines of the full table that belong to those rows (bench.py's latitude-band samples).  The two
 * searches of search_OverwrapRange are evaluated as written but only once per destination column (longitude
 * search: independent of jD) and once per destination row (latitude search: independent of iD). */
int orc_gen_jones99_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                         int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                         const double *y_LatIntWtS, const double *y_LatIntWtD,
                         int accuracy_order, int lon_mode, int jD0, int jD1, orc_table *out)
{
    const double PI = acos(-1.0);
    (void)y_LatD; (void)x_LonD;
    if (jD0 < 1) jD0 = 1;
    if (jD1 > nyd) jD1 = nyd;
    int *memo_rx = (int *)malloc(sizeof(int) * 2 * (size_t)(nxd + 1));
    for (int i = 0; i < 2 * (nxd + 1); i++) memo_rx[i] = -2;
    double *uS = (double *)malloc(sizeof(double) * (nxs + 1));
    double *uD = (double *)malloc(sizeof(double) * (nxd + 1));
    double *vS = (double *)malloc(sizeof(double) * (nys + 1));
    double *vD = (double *)malloc(sizeof(double) * (nyd + 1));
    calc_lon_edges(nxs, x_LonS, uS);
    calc_lon_edges(nxd, x_LonD, uD);
    calc_lat_edges(nys, y_LatIntWtS, vS);
    calc_lat_edges(nyd, y_LatIntWtD, vD);

    double *segLat = (double *)malloc(sizeof(double) * (nys + 2));
    double *w1 = (double *)malloc(sizeof(double) * (nys + 1));
    double *w2 = (double *)malloc(sizeof(double) * (nys + 1));
    int rc = 0;

    /* generalised longitude overlap (extension): per destination longitude a list of
     * (source longitude index, overlap fraction), west to east. */
    int general_lon = 0;
    int *glon_ptr = NULL, *glon_idx = NULL; double *glon_w = NULL;
    if (lon_mode == 1 && nxs != 1 && nxd != 1) {
        int same = (nxs == nxd);
        if (same) for (int i = 0; i < nxs; i++) if (x_LonS[i] != x_LonD[i]) { same = 0; break; }
        if (!same) {
            if (accuracy_order > 1) { rc = -3; goto done; }
            general_lon = 1;
            glon_ptr = (int *)malloc(sizeof(int) * (nxd + 1));
            int cap = 4 * (nxd + nxs) + 16, cnt = 0;
            glon_idx = (int *)malloc(sizeof(int) * cap);
            glon_w = (double *)malloc(sizeof(double) * cap);
            for (int id = 1; id <= nxd; id++) {
                double a = uD[id - 1], b = uD[id];
                glon_ptr[id - 1] = cnt;
                for (int shift = -1; shift <= 1; shift++) {
                    double off = 2.0 * PI * shift;
                    for (int m = 1; m <= nxs; m++) {
                        double lo = uS[m - 1] + off, hi = uS[m] + off;
                        double l = lo > a ? lo : a, h = hi < b ? hi : b;
                        double ov = h - l;
                        if (ov > 0.0) {
                            if (cnt == cap) {
                                cap *= 2;
                                glon_idx = (int *)realloc(glon_idx, sizeof(int) * cap);
                                glon_w = (double *)realloc(glon_w, sizeof(double) * cap);
                            }
                            glon_idx[cnt] = m; glon_w[cnt] = ov / (b - a); cnt++;
                        }
                    }
                }
            }
            glon_ptr[nxd] = cnt;
        }
    }

    for (int jD = jD0; jD <= jD1; jD++) {
        int memo_ry1 = -2, memo_ry2 = -2;
        for (int iD = 1; iD <= nxd; iD++) {
            /* ---- search_OverwrapRange (:384-440) ---- */
            int rx1 = -1, rx2 = -1, ry1 = -1, ry2 = -1;
            double X1 = uD[iD - 1], X2 = uD[iD], Y1 = vD[jD - 1], Y2 = vD[jD];
            if (general_lon) {
                rx1 = rx2 = 0;
            } else if (nxs == 1 && nxd != 1) {
                rx1 = 1; rx2 = 1;                                   /* :405-406 */
            } else if (nxs != 1 && nxd == 1) {
                rx1 = 1; rx2 = nxs;                                 /* :407-408 */
            } else if (memo_rx[2 * iD] != -2) {
                rx1 = memo_rx[2 * iD]; rx2 = memo_rx[2 * iD + 1];
            } else {
                for (int i = 1; i <= nxs; i++) {                    /* :410-418 */
                    if (uS[i - 1] <= X1 && X1 <= uS[i]) rx1 = i;
                    if (uS[i - 1] <= X2 && X2 <= uS[i]) { rx2 = i; break; }
                }
                if (rx1 < 0 || rx2 < 0) { rc = -1; goto done; }     /* unsupported by the reference */
                memo_rx[2 * iD] = rx1; memo_rx[2 * iD + 1] = rx2;
            }
            if (memo_ry1 != -2) {
                ry1 = memo_ry1; ry2 = memo_ry2;
            } else {
                for (int j = 1; j <= nys; j++) {                    /* :421-429 */
                    if (vS[j - 1] <= Y1 && Y1 <= vS[j]) ry1 = j;
                    if (vS[j - 1] <= Y2 && Y2 <= vS[j]) { ry2 = j; break; }
                }
                if (ry1 < 0 || ry2 < 0) { rc = -2; goto done; }     /* :433-438 "Exception.." stop */
                memo_ry1 = ry1; memo_ry2 = ry2;
            }

            /* ---- calc_RemappingWeight (:280-382) ---- */
            int nxr = general_lon ? (glon_ptr[iD] - glon_ptr[iD - 1]) : (rx2 - rx1 + 1);
            int nyr = ry2 - ry1 + 1;
            double DLon_k = 2.0 * PI / (double)nxd, DLon_nk, DLon_n;
            if (nxd == 1) { DLon_nk = 2.0 * PI / (double)nxs; DLon_n = DLon_nk; }
            else          { DLon_nk = 2.0 * PI / (double)nxd; DLon_n = 2.0 * PI; }
            segLat[0] = vD[jD - 1];
            for (int j = 1; j <= nyr - 1; j++) segLat[j] = vS[ry1 + j - 1];
            segLat[nyr] = vD[jD];
            double lat1_k = vD[jD - 1], lat2_k = vD[jD];
            double Ak = (sin(lat2_k) - sin(lat1_k)) * DLon_k;
            for (int j = 1; j <= nyr; j++) {
                double lat1_nk = segLat[j - 1], lat2_nk = segLat[j];
                if (general_lon) w1[j] = (sin(lat2_nk) - sin(lat1_nk)) / (sin(lat2_k) - sin(lat1_k));
                else             w1[j] = DLon_nk * (sin(lat2_nk) - sin(lat1_nk)) / Ak;          /* :347-350 */
            }
            if (accuracy_order > 1) {
                for (int j = 1; j <= nyr; j++) {                                                /* :354-365 */
                    double lat1_nk = segLat[j - 1], lat2_nk = segLat[j];
                    double lat1_n = vS[ry1 + j - 2], lat2_n = vS[ry1 + j - 1];
                    double An = (sin(lat2_n) - sin(lat1_n)) * DLon_n;
                    w2[j] = ((cos(lat2_nk) + lat2_nk * sin(lat2_nk))
                           - (cos(lat1_nk) + lat1_nk * sin(lat1_nk))) * DLon_nk / Ak
                          - ((cos(lat2_n) + lat2_n * sin(lat2_n))
                           - (cos(lat1_n) + lat1_n * sin(lat1_n))) * DLon_n * w1[j] / An;
                }
            }

            /* ---- emit (:240-272): m (lon) outer, n (lat) inner ---- */
            for (int m = 1; m <= nxr; m++) {
                for (int n = 1; n <= nyr; n++) {
                    int iS = general_lon ? glon_idx[glon_ptr[iD - 1] + m - 1] : rx1 + m - 1;
                    int jS = ry1 + n - 1;
                    double w = general_lon ? glon_w[glon_ptr[iD - 1] + m - 1] * w1[n] : w1[n];
                    if (fabs(w) > 1e-14) table_push(out, iD, jD, iS, jS, w);                    /* :245 */
                    if (accuracy_order > 1) {
                        int j1, j2;
                        if (jS == 1)        { j1 = jS;     j2 = jS + 1; }                       /* :252-258 */
                        else if (jS == nys) { j1 = jS - 1; j2 = jS; }
                        else                { j1 = jS - 1; j2 = jS + 1; }
                        double DLat = y_LatS[j2 - 1] - y_LatS[j1 - 1];
                        if (fabs(w2[n]) > 1e-14) {                                              /* :262-267 */
                            table_push(out, iD, jD, iS, j1, -w2[n] / DLat);
                            table_push(out, iD, jD, iS, j2, +w2[n] / DLat);
                        }
                    }
                }
            }
        }
    }
done:
    free(uS); free(uD); free(vS); free(vD); free(segLat); free(w1); free(w2);
    free(glon_ptr); free(glon_idx); free(glon_w); free(memo_rx);
    return rc;
}

int orc_gen_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                    int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                    const double *y_LatIntWtS, const double *y_LatIntWtD,
                    int accuracy_order, int lon_mode, orc_table *out)
{
    return orc_gen_jones99_rows(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                                accuracy_order, lon_mode, 1, nyd, out);
}

/* ------------------------------------------------------------------ bilinear generator */

/* ref common/grid_mapping_util.f90:32-177 */
int orc_gen_bilinear_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                          int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                          int lon_mode, int jr0, int jr1, orc_table *out)
{
    const double PI = acos(-1.0);
    double dlon_r = 360.0 / (double)nxr;     /* :78 */
    double dlon_s = 360.0 / (double)nxs;     /* :79 */
    if (jr0 < 1) jr0 = 1;
    if (jr1 > nyr) jr1 = nyr;
    for (int jr = jr0; jr <= jr1; jr++) {     /* the reference: 1..nyr; a sub-range = those lines of the full table */
        /* get_correspondID_latS (:130-152) */
        double latR = y_LatR[jr - 1];
        int js = -1, extp = 1;
        for (int j = 1; j <= nys - 1; j++) {
            if (y_LatS[j - 1] < latR && latR <= y_LatS[j]) { js = j; extp = 0; break; }
        }
        if (extp) {
            if (latR <= y_LatS[0]) js = 1;
            if (latR > y_LatS[nys - 1]) js = nys;
        }
        if (nxr == 1) {                                               /* :88-100 */
            if (!extp) {
                double b1 = (latR - y_LatS[js - 1]) / (y_LatS[js] - y_LatS[js - 1]);   /* cal_coef_axisym :167-175 */
                double c1 = (1.0 - b1) / nxs, c2 = b1 / nxs;
                for (int is = 1; is <= nxs; is++) {
                    table_push(out, 1, jr, is, js, c1);
                    table_push(out, 1, jr, is, js % nys + 1, c2);
                }
            } else {
                for (int is = 1; is <= nxs; is++) table_push(out, 1, jr, is, js, 1.0 / nxs);
            }
        } else if (nxs == 1) {                                        /* :101-112 */
            if (!extp) {
                double b1 = (latR - y_LatS[js - 1]) / (y_LatS[js] - y_LatS[js - 1]);
                double c1 = 1.0 - b1, c2 = b1;
                for (int ir = 1; ir <= nxr; ir++) {
                    table_push(out, ir, jr, 1, js, c1);
                    table_push(out, ir, jr, 1, js % nys + 1, c2);
                }
            } else {
                for (int ir = 1; ir <= nxr; ir++) table_push(out, ir, jr, 1, js, 1.0);
            }
        } else {                                                      /* :113-125; extp_flag is NOT tested here */
            int jlo = js, jhi = js + 1;
            if (js == nys) { jlo = nys - 1; jhi = nys; }              /* defect A5-1, see file header */
            int jhi_out = (js == nys) ? jhi : (js % nys + 1);         /* mod(js,nys)+1 */
            for (int ir = 1; ir <= nxr; ir++) {
                int is = (int)(dlon_r * (ir - 1) / dlon_s) + 1;       /* :116 */
                int is2 = is % nxs + 1;
                double xc = x_LonR[ir - 1], yc = latR;
                double x1 = x_LonS[is - 1], y1 = y_LatS[jlo - 1];
                double x3 = x_LonS[is2 - 1], y3 = y_LatS[jhi - 1];
                if (lon_mode == 1 && x3 <= x1) x3 += 2.0 * PI;        /* extension: unwrap east neighbour */
                double a1 = (xc - x1) / (x3 - x1), a2 = 1.0 - a1;     /* cal_coef :154-165 */
                double b1 = (yc - y1) / (y3 - y1), b2 = 1.0 - b1;
                table_push(out, ir, jr, is,  jlo,     a2 * b2);
                table_push(out, ir, jr, is2, jlo,     a1 * b2);
                table_push(out, ir, jr, is2, jhi_out, a1 * b1);
                table_push(out, ir, jr, is,  jhi_out, a2 * b1);
            }
        }
    }
    return 0;
}

int orc_gen_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                     int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                     int lon_mode, orc_table *out)
{
    return orc_gen_bilinear_rows(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, 1, nyr, out);
}

/* ------------------------------------------------------------------ exchange grid */

/* ref common/cal_mappingtable.f90:10-49 (make_mapping_table) with cal_coef :59-76: the older stand-alone
 * bilinear generator on REGULAR grids in degrees -- longitudes 360/nx apart starting at 0, latitudes
 * 180/(ny-1) apart from pole to pole.  An entry is written only when its coefficient is > 0 (:40-43), in the
 * order coef(1..4) = (a2*b2, a1*b2, a1*b1, a2*b1); the east / north neighbours wrap with mod (:41-43).
 * The program around it (:81-109) refers to undefined names and is not built by the reference. */
int orc_make_mapping_table(int nx_r, int ny_r, int nx_s, int ny_s, orc_table *out)
{
    if (nx_r < 1 || nx_s < 1 || ny_r < 2 || ny_s < 2) return 1;
    double dx_r = 360.0 / nx_r;                 /* :20 */
    double dy_r = 180.0 / (ny_r - 1);           /* :21 */
    double dx_s = 360.0 / nx_s;                 /* :22 */
    double dy_s = 180.0 / (ny_s - 1);           /* :23 */
    for (int j = 1; j <= ny_r; j++) {           /* :25 */
        double yr = dy_r * (j - 1);
        int js = (int)(yr / dy_s) + 1;          /* :27 int() truncates */
        double ys1 = (js - 1) * dy_s;
        double ys3 = ys1 + dy_s;
        for (int i = 1; i <= nx_r; i++) {       /* :31 */
            double xr = dx_r * (i - 1);
            int is = (int)(xr / dx_s) + 1;      /* :34 */
            double xs1 = (is - 1) * dx_s;
            double xs3 = xs1 + dx_s;
            /* cal_coef(xr, yr, xs1, ys1, xs3, ys3, coef) :66-74 */
            double alpha1 = (xr - xs1) / (xs3 - xs1);
            double alpha2 = 1.0 - alpha1;
            double beta1 = (yr - ys1) / (ys3 - ys1);
            double beta2 = 1.0 - beta1;
            double c1 = alpha2 * beta2, c2 = alpha1 * beta2, c3 = alpha1 * beta1, c4 = alpha2 * beta1;
            if (c1 > 0.0) table_push(out, i, j, is, js, c1);                               /* :40 */
            if (c2 > 0.0) table_push(out, i, j, is % nx_s + 1, js, c2);                    /* :41 */
            if (c3 > 0.0) table_push(out, i, j, is % nx_s + 1, js % ny_s + 1, c3);         /* :42 */
            if (c4 > 0.0) table_push(out, i, j, is, js % ny_s + 1, c4);                    /* :43 */
        }
    }
    return 0;
}

/* ref tool/gmapgen/gmapgen_main.f90:336-405 (generate_surface_exchage_grid) + sort :407-426.
 * Longitudes of the exchange grid are those of the atmosphere (:349-352) and are not returned.
 * y_LatS / y_IntWtLatS must have room for jma + jmo - 1 entries. */
int orc_exchange_grid(int jma, const double *y_LatA, const double *y_IntWtLatA,
                      int jmo, const double *y_IntWtLatO,
                      int *jms, double *y_LatS, double *y_IntWtLatS)
{
    const double PI = acos(-1.0);
    if (jma == jmo) {                                                  /* :354-359 */
        *jms = jmo;
        for (int j = 0; j < jmo; j++) { y_LatS[j] = y_LatA[j]; y_IntWtLatS[j] = y_IntWtLatA[j]; }
        return 0;
    }
    int jms_ = jmo + jma - 1;                                          /* :361 */
    double *fja = (double *)malloc(sizeof(double) * (jma + 1));
    double *fjo = (double *)malloc(sizeof(double) * (jmo + 1));
    double *fjs = (double *)malloc(sizeof(double) * (jms_ + 1));
    fja[0] = -PI / 2.0;                                                /* :364-368 */
    for (int j = 1; j <= jma - 1; j++) fja[j] = asin(y_IntWtLatA[j - 1] + sin(fja[j - 1]));
    fja[jma] = PI / 2.0;
    fjo[0] = -PI / 2.0;                                                /* :370-374 */
    for (int j = 1; j <= jmo - 1; j++) fjo[j] = asin(y_IntWtLatO[j - 1] + sin(fjo[j - 1]));
    fjo[jmo] = PI / 2.0;
    for (int j = 0; j <= jma - 1; j++) fjs[j] = fja[j];                /* :376 */
    for (int j = jma; j <= jms_; j++) fjs[j] = fjo[j - jma + 1];       /* :377 */
    int N = jms_ + 1;                                                  /* sort :407-426 */
    for (int i = 0; i < N - 1; i++)
        for (int j = i + 1; j < N; j++)
            if (fjs[i] > fjs[j]) { double t = fjs[i]; fjs[i] = fjs[j]; fjs[j] = t; }
    int n = 0;                                                         /* :383-392 */
    for (int j = 1; j <= jms_; j++) {
        double intWt = sin(fjs[j]) - sin(fjs[j - 1]);
        if (fabs(intWt) > 1e-12) {
            y_LatS[n] = 0.5 * (fjs[j - 1] + fjs[j]);
            y_IntWtLatS[n] = intWt;
            n++;
        }
    }
    *jms = n;
    free(fja); free(fjo); free(fjs);
    return 0;
}

/* ------------------------------------------------------------------ table file I/O */

/* One entry per line, "iD jD iS jS coef", list-directed compatible
 * (ref common/grid_mapping_util_jones99.f90:246-247). 17 significant digits so the
 * text round trip is exact. */
int orc_table_write_text(const orc_table *t, const char *filename)
{
    FILE *f = fopen(filename, "w");
    if (!f) return -1;
    for (int64_t k = 0; k < t->n; k++)
        fprintf(f, "%12d%12d%12d%12d  %24.16E\n", t->iD[k], t->jD[k], t->iS[k], t->jS[k], t->coef[k]);
    fclose(f);
    return 0;
}

/* ref common/grid_mapping_util_jones99.f90:479-504: read(deviceID,*) ir, jr, is, js, coef */
int orc_table_read_text(const char *filename, orc_table *out)
{
    FILE *f = fopen(filename, "r");
    if (!f) return -1;
    char line[512];
    while (fgets(line, sizeof line, f)) {
        for (char *p = line; *p; p++) { if (*p == ',') *p = ' '; if (*p == 'D' || *p == 'd') *p = 'E'; }
        char *p = line, *e;
        long v[4]; int ok = 1;
        for (int k = 0; k < 4; k++) { v[k] = strtol(p, &e, 10); if (e == p) { ok = 0; break; } p = e; }
        if (!ok) continue;
        double c = strtod(p, &e);
        if (e == p) continue;
        table_push(out, (int)v[0], (int)v[1], (int)v[2], (int)v[3], c);
    }
    fclose(f);
    return 0;
}

/* ref :498-500: recv_index = ir + GNXR*(jr-1); send_index = is + GNXS*(js-1) */
void orc_table_to_index(const orc_table *t, int gnxs, int gnxr,
                        int32_t *send_index, int32_t *recv_index, double *coef_s)
{
    for (int64_t n = 0; n < t->n; n++) {
        recv_index[n] = t->iD[n] + gnxr * (t->jD[n] - 1);
        send_index[n] = t->iS[n] + gnxs * (t->jS[n] - 1);
        coef_s[n] = t->coef[n];
    }
}

/* ------------------------------------------------------------------ remap apply */

/* ref common/interpolation_data_latlon_mod.f90:293-302.  Serial, d outer / op inner,
 * zero-fill of ALL of recv_data(rn1,rn2) first. Indices are 1-based. */
void orc_remap_apply(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                     const double *coef, const double *send, int sn1, int sn2,
                     double *recv, int rn1, int rn2, int num_of_data)
{
    (void)sn2;
    for (int64_t k = 0; k < (int64_t)rn1 * rn2; k++) recv[k] = 0.0;              /* :293 */
    for (int d = 0; d < num_of_data; d++) {                                      /* :295 */
        const double *s = send + (int64_t)d * sn1;
        double *r = recv + (int64_t)d * rn1;
        for (int64_t i = 0; i < nops; i++) {                                     /* :296 */
            int sp = send_index[i] - 1, rp = recv_index[i] - 1;
            r[rp] = r[rp] + s[sp] * coef[i];                                     /* :299-300 */
        }
    }
}

/* ------------------------------------------------------------------ bulk flux */

/* constants: ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:35-58 */
static const double FKarm = 0.4;
static const double GasRUniv = 8.3144621;
static const double StB = 5.670373e-8;
static const double Grav_sfc = 9.8;
static const double MolWtDry = 1.8e-2;
static const double MolWtWet = 1.8e-2;
static const double CpDry_sfc = 1616.0;
static const double LatentHeat = 2425300.0;
static const double LatentHeatFusion = 334000.0;
static const double RefPress = 1e5;
static const double Es0 = 611.0;
static const double RoughLength = 1e-4;
static const double RoughLenHeatFactor = 1.0;
/* limits: ref :574-591 (read_config) */
static const double VelMinForRi = 0.01, VelMinForVel = 0.01, VelMinForTemp = 0.01, VelMinForQVap = 0.01;
static const double VelMaxForVel = 1000.0, VelMaxForTemp = 1000.0, VelMaxForQVap = 1000.0;
static const double VelBulkCoefMin = 0.0, TempBulkCoefMin = 0.0, QVapBulkCoefMin = 0.0;
static const double VelBulkCoefMax = 1.0, TempBulkCoefMax = 1.0, QVapBulkCoefMax = 1.0;

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

/* ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:443-572 (BulkCoefL82) */
static void BulkCoefL82(int IA, int JA,
    const double *Ri, const double *z0m, const double *z0h,
    const double *SfcHeight, const double *Height,
    const double *CMn, const double *CHn, const char *CalcFlag,
    double *CM, double *CH, double *CQ)
{
    int IS = 1, IE = IA - 2, JS = 1, JE = JA - 2;
    #pragma omp parallel for collapse(2)
    for (int j = JS; j <= JE; j++)
    for (int i = IS; i <= IE; i++) {
        int c = i + IA * j;
        if (CalcFlag[c]) {
            if (Ri[c] > 0.0) {                                                    /* :490-506 */
                CM[c] = CMn[c] / (1.0 + 10.0 * Ri[c] / sqrt(1.0 + 5.0 * Ri[c]));
                CH[c] = CHn[c] / (1.0 + 15.0 * Ri[c] * sqrt(1.0 + 5.0 * Ri[c]));
                CQ[c] = CH[c];
            } else {                                                              /* :510-534 */
                CM[c] = CMn[c] * (1.0 - 10.0 * Ri[c]
                        / (1.0 + 75.0 * CMn[c]
                           * sqrt(-(Height[c] - SfcHeight[c] + z0m[c]) / z0m[c] * Ri[c])));
                CH[c] = CHn[c] * (1.0 - 15.0 * Ri[c]
                        / (1.0 + 75.0 * CHn[c]
                           * sqrt(-(Height[c] - SfcHeight[c] + z0h[c]) / z0h[c] * Ri[c])));
                CQ[c] = CH[c];
            }
        } else {                                                                  /* :536-540 */
            CM[c] = 0.0; CH[c] = 0.0; CQ[c] = 0.0;
        }
        /* Monin-Obukhov length (:545-550): unused local, not computed (B-3) */
        CM[c] = dmax(dmin(CM[c], VelBulkCoefMax), VelBulkCoefMin);               /* :555-565 */
        CH[c] = dmax(dmin(CH[c], TempBulkCoefMax), TempBulkCoefMin);
        CQ[c] = dmax(dmin(CQ[c], QVapBulkCoefMax), QVapBulkCoefMin);
    }
}

/* ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:108-439 (DSFCM_Util_SfcBulkFlux_Get).
 * Arrays are Fortran column-major (IA,JA[,n]); halo 1, interior IS:IE,JS:JE
 * (sfc/DSFCM_Admin_Grid_mod.f90:39-50). Multi-pass with full-grid temporaries, as the reference. */
void orc_bulkflux(int IA, int JA,
    double *WSX, double *WSY, double *SenH, double *QVapM, double *LatH,
    double *VelTC, double *TempTC, double *QVapTC,
    double *Del,
    double *SUw, double *LUw,
    double *HFns, double *HFsr, double *DHFDTs,
    const double *WindU, const double *WindV, const double *SfcAirTemp, const double *QVap1,
    const double *SDw, const double *LDw,
    const double *Coef1, const double *Coef2,
    double *SfcTemp, double *SfcAlbedo, const double *SIceCon,
    const double *Sig1Info, const double *SfcHeight, const double *SfcPress)
{
    const int SPMAX = 3;
    const int IS = 1, IE = IA - 2, JS = 1, JE = JA - 2;   /* 0-based interior */
    const size_t N2 = (size_t)IA * JA;
    const double GasRDry = GasRUniv / MolWtDry;           /* :49-51 */
    const double GasRWet = GasRUniv / MolWtWet;
    const double EpsV = MolWtWet / MolWtDry;
#define A3(a, c, n) a[(c) + N2 * (size_t)(n)]

    double *z0m = (double *)malloc(sizeof(double) * N2 * 2);
    double *z0h = (double *)malloc(sizeof(double) * N2 * 2);
    double *HumdCoef = (double *)malloc(sizeof(double) * N2 * 3);
    double *RiNum = (double *)malloc(sizeof(double) * N2 * 2);
    double *CMn = (double *)malloc(sizeof(double) * N2);
    double *CHn = (double *)malloc(sizeof(double) * N2);
    double *CM = (double *)malloc(sizeof(double) * N2 * 2);
    double *CH = (double *)malloc(sizeof(double) * N2 * 2);
    double *CQ = (double *)malloc(sizeof(double) * N2 * 2);
    double *VelAbs = (double *)malloc(sizeof(double) * N2);
    double *VirTemp = (double *)malloc(sizeof(double) * N2);
    double *SfcVirTemp = (double *)malloc(sizeof(double) * N2 * 2);
    double *SfcQVapSat = (double *)malloc(sizeof(double) * N2 * 2);
    double *Exner = (double *)malloc(sizeof(double) * N2);
    double *SfcExner = (double *)malloc(sizeof(double) * N2);
    double *Frac = (double *)malloc(sizeof(double) * N2 * 2);
    double *Height = (double *)malloc(sizeof(double) * N2);
    double *Press1 = (double *)malloc(sizeof(double) * N2);
    char *CalcFlag = (char *)malloc(N2);
    double LatentHeatLocal[2];

    for (size_t k = 0; k < N2 * 2; k++) { z0m[k] = RoughLength; }                 /* :194-196 */
    for (size_t k = 0; k < N2 * 2; k++) { z0h[k] = RoughLenHeatFactor * z0m[k]; }
    for (size_t k = 0; k < N2 * 3; k++) { HumdCoef[k] = 1.0; }
    LatentHeatLocal[0] = LatentHeat;                                              /* :198-199 */
    LatentHeatLocal[1] = LatentHeat + LatentHeatFusion;

    #pragma omp parallel for collapse(2)
    for (int j = JS; j <= JE; j++)                                                /* :203-242 */
    for (int i = IS; i <= IE; i++) {
        size_t c = i + (size_t)IA * j;
        A3(Frac, c, 0) = 1.0 - SIceCon[c];
        A3(Frac, c, 1) = SIceCon[c];
        for (int n = 0; n < SPMAX - 1; n++) {
            A3(SfcQVapSat, c, n) = EpsV * Es0 / SfcPress[c]
                * m_exp(LatentHeatLocal[n] / GasRWet * (1.0 / 273.0 - 1.0 / A3(SfcTemp, c, n)));
            A3(SfcVirTemp, c, n) = A3(SfcTemp, c, n) * (1.0 + (((1.0 / EpsV) - 1.0) * A3(SfcQVapSat, c, n)));
        }
        VirTemp[c] = SfcAirTemp[c] * (1.0 + (((1.0 / EpsV) - 1.0) * QVap1[c]));
        Press1[c] = SfcPress[c] * Sig1Info[0];
        Exner[c] = m_pow(Press1[c] / RefPress, GasRDry / CpDry_sfc);
        SfcExner[c] = m_pow(SfcPress[c] / RefPress, GasRDry / CpDry_sfc);
        VelAbs[c] = sqrt(WindU[c] * WindU[c] + WindV[c] * WindV[c]);
        Height[c] = SfcHeight[c] + GasRDry / Grav_sfc * VirTemp[c] * (1.0 - Sig1Info[0]);
        A3(WSX, c, 2) = 0.0; A3(WSY, c, 2) = 0.0; A3(SenH, c, 2) = 0.0; A3(LatH, c, 2) = 0.0;
        A3(QVapM, c, 2) = 0.0; A3(SUw, c, 2) = 0.0; A3(LUw, c, 2) = 0.0;
        A3(SfcTemp, c, 2) = 0.0; A3(SfcAlbedo, c, 2) = 0.0;
        A3(VelTC, c, 2) = 0.0; A3(TempTC, c, 2) = 0.0; A3(QVapTC, c, 2) = 0.0;
    }

    for (int n = 0; n < SPMAX - 1; n++) {                                         /* :244 */
        #pragma omp parallel for collapse(2)
        for (int j = JS; j <= JE; j++)                                            /* :247-276 */
        for (int i = IS; i <= IE; i++) {
            size_t c = i + (size_t)IA * j;
            double tmp = FKarm / m_log((Height[c] - SfcHeight[c] + A3(z0m, c, n)) / A3(z0m, c, n));
            CMn[c] = tmp * tmp;
            CHn[c] = tmp * (FKarm / m_log((Height[c] - SfcHeight[c] + A3(z0h, c, n)) / A3(z0h, c, n)));
            double vr = dmax(VelAbs[c], VelMinForRi);
            A3(RiNum, c, n) = Grav_sfc / (A3(SfcVirTemp, c, n) / SfcExner[c])
                * (VirTemp[c] / Exner[c] - A3(SfcVirTemp, c, n) / SfcExner[c])
                / (vr * vr)
                * (Height[c] - SfcHeight[c]);
            if (n == 0) CalcFlag[c] = 1;
            else        CalcFlag[c] = (A3(Frac, c, n) > 1e-12);
        }

        BulkCoefL82(IA, JA, RiNum + N2 * n, z0m + N2 * n, z0h + N2 * n, SfcHeight, Height,
                    CMn, CHn, CalcFlag, CM + N2 * n, CH + N2 * n, CQ + N2 * n);   /* :278-284 */

        #pragma omp parallel for collapse(2)
        for (int j = JS; j <= JE; j++)                                            /* :286-349 */
        for (int i = IS; i <= IE; i++) {
            size_t c = i + (size_t)IA * j;
            A3(VelTC, c, n) = A3(CM, c, n) * SfcPress[c] / (GasRDry * A3(SfcVirTemp, c, n))
                * dmin(dmax(VelAbs[c], VelMinForVel), VelMaxForVel);
            A3(TempTC, c, n) = A3(CH, c, n) * SfcPress[c] / (GasRDry * A3(SfcVirTemp, c, n))
                * dmin(dmax(VelAbs[c], VelMinForTemp), VelMaxForTemp);
            A3(QVapTC, c, n) = A3(CQ, c, n) * SfcPress[c] / (GasRDry * A3(SfcVirTemp, c, n))
                * dmin(dmax(VelAbs[c], VelMinForQVap), VelMaxForQVap);
            if (CalcFlag[c]) {
                A3(WSX, c, n) = -A3(VelTC, c, n) * WindU[c];
                A3(WSY, c, n) = -A3(VelTC, c, n) * WindV[c];
                A3(SenH, c, n) = -CpDry_sfc * SfcExner[c] * A3(TempTC, c, n)
                    * (SfcAirTemp[c] / Exner[c] - A3(SfcTemp, c, n) / SfcExner[c]);
                A3(QVapM, c, n) = -A3(HumdCoef, c, n) * A3(QVapTC, c, n)
                    * (QVap1[c] - A3(SfcQVapSat, c, n));
                A3(LatH, c, n) = LatentHeatLocal[n] * A3(QVapM, c, n);
                {
                    double t = A3(SfcTemp, c, n), t2 = t * t;
                    A3(LUw, c, n) = StB * (t2 * t2);                  /* T**4 */
                }
                A3(SUw, c, n) = A3(SfcAlbedo, c, n) * SDw[c];
                {
                    double t = A3(SfcTemp, c, n), t2 = t * t;
                    A3(SfcTemp, c, 2) = A3(SfcTemp, c, 2) + A3(Frac, c, n) * (t2 * t2);
                }
                A3(SfcAlbedo, c, 2) = A3(SfcAlbedo, c, 2) + A3(Frac, c, n) * A3(SfcAlbedo, c, n);
                A3(WSX, c, 2) = A3(WSX, c, 2) + A3(Frac, c, n) * A3(WSX, c, n);
                A3(WSY, c, 2) = A3(WSY, c, 2) + A3(Frac, c, n) * A3(WSY, c, n);
                A3(SenH, c, 2) = A3(SenH, c, 2) + A3(Frac, c, n) * A3(SenH, c, n);
                A3(QVapM, c, 2) = A3(QVapM, c, 2) + A3(Frac, c, n) * A3(QVapM, c, n);
                A3(LatH, c, 2) = A3(LatH, c, 2) + A3(Frac, c, n) * A3(LatH, c, n);
                A3(LUw, c, 2) = A3(LUw, c, 2) + A3(Frac, c, n) * A3(LUw, c, n);
                A3(SUw, c, 2) = A3(SUw, c, 2) + A3(Frac, c, n) * A3(SUw, c, n);
                A3(VelTC, c, 2) = A3(VelTC, c, 2) + A3(Frac, c, n) * A3(VelTC, c, n);
                A3(TempTC, c, 2) = A3(TempTC, c, 2) + A3(Frac, c, n) * A3(TempTC, c, n);
                A3(QVapTC, c, 2) = A3(QVapTC, c, 2) + A3(Frac, c, n) * A3(QVapTC, c, n);
            } else {
                A3(WSX, c, n) = 0.0; A3(WSY, c, n) = 0.0; A3(SenH, c, n) = 0.0; A3(QVapM, c, n) = 0.0;
                A3(LatH, c, n) = 0.0; A3(LUw, c, n) = 0.0; A3(SUw, c, n) = 0.0;
            }
        }
    }

    #pragma omp parallel for collapse(2)
    for (int j = JS; j <= JE; j++)                                                /* :353-382 */
    for (int i = IS; i <= IE; i++) {
        size_t c = i + (size_t)IA * j;
        double a_Gamma[4], a_f[4];
        double DFsDT1 = -CpDry_sfc * SfcExner[c] * A3(TempTC, c, 2) / Exner[c];
        a_Gamma[0] = 1.0 / (A3(Coef1, c, 0) + A3(VelTC, c, 2));
        a_Gamma[1] = 1.0 / (A3(Coef1, c, 1) + A3(VelTC, c, 2));
        a_Gamma[2] = 1.0 / (A3(Coef1, c, 2) - DFsDT1);
        a_Gamma[3] = 1.0 / (A3(Coef1, c, 3) + A3(HumdCoef, c, 2) * A3(QVapTC, c, 2));
        a_f[0] = a_Gamma[0] * (A3(WSX, c, 2) + A3(Coef2, c, 0));
        a_f[1] = a_Gamma[1] * (A3(WSY, c, 2) + A3(Coef2, c, 1));
        a_f[2] = a_Gamma[2] * (A3(SenH, c, 2) + A3(Coef2, c, 2));
        a_f[3] = a_Gamma[3] * (A3(QVapM, c, 2) + A3(Coef2, c, 3));
        for (int k = 0; k < 4; k++) A3(Del, c, k) = a_f[k];
        double lat3 = 0.0;
        for (int n = 0; n < SPMAX; n++) {
            A3(WSX, c, n) = A3(WSX, c, n) - A3(VelTC, c, n) * a_f[0];
            A3(WSY, c, n) = A3(WSY, c, n) - A3(VelTC, c, n) * a_f[1];
            A3(SenH, c, n) = A3(SenH, c, n)
                - CpDry_sfc * SfcExner[c] / Exner[c] * A3(TempTC, c, n) * a_f[2];
            A3(QVapM, c, n) = A3(QVapM, c, n) - A3(HumdCoef, c, n) * A3(QVapTC, c, n) * a_f[3];
            if (n < 2) {
                A3(LatH, c, n) = LatentHeatLocal[n] * A3(QVapM, c, n);
                lat3 = lat3 + A3(Frac, c, n) * A3(LatH, c, n);
            } else {
                A3(LatH, c, n) = lat3;                                 /* defect B-1, see file header */
            }
        }
    }

    for (int n = 0; n < SPMAX - 1; n++) {                                         /* :384-415 */
        #pragma omp parallel for collapse(2)
        for (int j = JS; j <= JE; j++)
        for (int i = IS; i <= IE; i++) {
            size_t c = i + (size_t)IA * j;
            int flag = (n == 0) ? 1 : (A3(Frac, c, n) > 1e-12);
            if (flag) {
                A3(HFns, c, n) = +A3(LUw, c, n) - LDw[c] + A3(LatH, c, n) + A3(SenH, c, n);
                A3(HFsr, c, n) = A3(SUw, c, n) - SDw[c];
                double t = A3(SfcTemp, c, n);
                A3(DHFDTs, c, n) = +4.0 * StB * (t * t * t)
                    + CpDry_sfc * A3(TempTC, c, n)
                    + LatentHeatLocal[n] * A3(HumdCoef, c, n) * A3(QVapTC, c, n)
                      * (LatentHeatLocal[n] * A3(SfcQVapSat, c, n) / (GasRWet * (t * t)));
            } else {
                A3(HFns, c, n) = 0.0; A3(HFsr, c, n) = 0.0; A3(DHFDTs, c, n) = 0.0;
            }
        }
    }
#undef A3
    free(z0m); free(z0h); free(HumdCoef); free(RiNum); free(CMn); free(CHn);
    free(CM); free(CH); free(CQ); free(VelAbs); free(VirTemp); free(SfcVirTemp);
    free(SfcQVapSat); free(Exner); free(SfcExner); free(Frac); free(Height); free(Press1);
    free(CalcFlag);
}

/* ------------------------------------------------------------------ implicit coupling */

/* ref atm/dcpam_sfc_implicit_coupling_mod.f90:420-426 (Init: module save matrices :16-18) */
orc_vdiff *orc_vdiff_new(int imax, int jmax, int kmax, int ncmax, int index_h2ovap,
                         double Grav, double CpDry, double GasRDry, double DelTime)
{
    orc_vdiff *h = (orc_vdiff *)calloc(1, sizeof(orc_vdiff));
    h->imax = imax; h->jmax = jmax; h->kmax = kmax; h->ncmax = ncmax; h->index_h2ovap = index_h2ovap;
    h->Grav = Grav; h->CpDry = CpDry; h->GasRDry = GasRDry; h->DelTime = DelTime;
    size_t n = (size_t)imax * jmax * kmax * 3;
    h->UVMtx = (double *)malloc(sizeof(double) * n);
    h->TempMtx = (double *)malloc(sizeof(double) * n);
    h->QMixMtx = (double *)malloc(sizeof(double) * n);
    return h;
}

void orc_vdiff_free(orc_vdiff *h)
{
    if (!h) return;
    free(h->UVMtx); free(h->TempMtx); free(h->QMixMtx); free(h);
}

/* matrix element (col, k=1..K, d=-1..1) */
#define MTX(M, c, k, d) M[(c) + NC * ((size_t)((k) - 1) + (size_t)K * ((d) + 1))]
/* full-level array (col, k=1..K) and half-level array (col, k=0..K) */
#define ZL(a, c, k) a[(c) + NC * (size_t)((k) - 1)]
#define RL(a, c, k) a[(c) + NC * (size_t)(k)]

/* ref :380-402 (Solve_TriDiagSystem_Forward); stops at k=2 (defect C-1). */
static void tridiag_forward(size_t NC, int K, double *Mtx, double *RHS)
{
    int k = K;
    for (size_t c = 0; c < NC; c++) {                                             /* :388-391 */
        MTX(Mtx, c, k, 0) = MTX(Mtx, c, k, 0) / MTX(Mtx, c, k, -1);
        ZL(RHS, c, k) = ZL(RHS, c, k) / MTX(Mtx, c, k, -1);
        MTX(Mtx, c, k, -1) = 1.0;
    }
    for (k = K - 1; k >= 2; k--) {                                                /* :393-400 */
        for (size_t c = 0; c < NC; c++) {
            double den = MTX(Mtx, c, k, -1) * MTX(Mtx, c, k + 1, 0);
            MTX(Mtx, c, k, 0) = (MTX(Mtx, c, k, 0) * MTX(Mtx, c, k + 1, 0) - MTX(Mtx, c, k, 1)) / den;
            ZL(RHS, c, k) = (ZL(RHS, c, k) * MTX(Mtx, c, k + 1, 0) - MTX(Mtx, c, k, 1) * ZL(RHS, c, k + 1)) / den;
            MTX(Mtx, c, k, -1) = 1.0;
            MTX(Mtx, c, k, 1) = 0.0;
        }
    }
}

/* ref :404-418 (Solve_TriDiagSystem_Backward) */
static void tridiag_backward(size_t NC, int K, const double *Mtx, double *RHS)
{
    for (int k = 2; k <= K; k++)
        for (size_t c = 0; c < NC; c++)
            ZL(RHS, c, k) = (ZL(RHS, c, k) - MTX(Mtx, c, k, -1) * ZL(RHS, c, k - 1)) / MTX(Mtx, c, k, 0);
}

/* ref :72-378 (SfcImplicitCoupling_VDiffForward).  Arrays are (0:imax-1,1:jmax,level[,n]),
 * column index fastest, level slowest. */
void orc_vdiff_forward(orc_vdiff *h,
    const double *MomFluxX, const double *MomFluxY, const double *HeatFlux, const double *QMixFlux,
    const double *Press, const double *zExner, const double *rExner,
    const double *VirTemp, const double *Height,
    const double *VelDiffCoef, const double *TempDiffCoef, const double *QMixDiffCoef,
    double *DUDt, double *DVDt, double *DTempDt, double *DQMixDt,
    double *Coef1, double *Coef2)
{
    const size_t NC = (size_t)h->imax * h->jmax;
    const int K = h->kmax, ncmax = h->ncmax, iq = h->index_h2ovap - 1;
    const double Grav = h->Grav, CpDry = h->CpDry, GasRDry = h->GasRDry, DelTime = h->DelTime;
    double *UV = h->UVMtx, *TM = h->TempMtx, *QM = h->QMixMtx;
    double *VelTC = (double *)malloc(sizeof(double) * NC * (K + 1));
    double *TempTC = (double *)malloc(sizeof(double) * NC * (K + 1));
    double *QMixTC = (double *)malloc(sizeof(double) * NC * (K + 1));
    double *Tmp = (double *)malloc(sizeof(double) * NC);
    double *Save = (double *)malloc(sizeof(double) * NC * K * 3);

    for (size_t c = 0; c < NC; c++) {                                             /* :189-194 */
        RL(VelTC, c, 0) = 0.0; RL(VelTC, c, K) = 0.0;
        RL(TempTC, c, 0) = 0.0; RL(TempTC, c, K) = 0.0;
        RL(QMixTC, c, 0) = 0.0; RL(QMixTC, c, K) = 0.0;
    }
    for (int k = 1; k <= K - 1; k++) {                                            /* :196-203 */
        for (size_t c = 0; c < NC; c++)
            Tmp[c] = RL(Press, c, k) / (GasRDry * RL(VirTemp, c, k)) / (ZL(Height, c, k + 1) - ZL(Height, c, k));
        for (size_t c = 0; c < NC; c++) RL(VelTC, c, k) = RL(VelDiffCoef, c, k) * Tmp[c];
        for (size_t c = 0; c < NC; c++) RL(TempTC, c, k) = RL(TempDiffCoef, c, k) * Tmp[c];
        for (size_t c = 0; c < NC; c++) RL(QMixTC, c, k) = RL(QMixDiffCoef, c, k) * Tmp[c];
    }

    /* UV matrix :207-233, QMix matrix :265-293 (same form) */
    for (int pass = 0; pass < 2; pass++) {
        double *M = pass == 0 ? UV : QM;
        const double *T = pass == 0 ? VelTC : QMixTC;
        int k = 1;
        for (size_t c = 0; c < NC; c++) {
            MTX(M, c, k, -1) = 0.0;
            MTX(M, c, k, 0) = -(RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime) + RL(T, c, k);
            MTX(M, c, k, 1) = -RL(T, c, k);
        }
        for (k = 2; k <= K - 1; k++)
            for (size_t c = 0; c < NC; c++) {
                MTX(M, c, k, -1) = -RL(T, c, k - 1);
                MTX(M, c, k, 0) = -(RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime)
                                  + RL(T, c, k - 1) + RL(T, c, k);
                MTX(M, c, k, 1) = -RL(T, c, k);
            }
        k = K;
        for (size_t c = 0; c < NC; c++) {
            MTX(M, c, k, -1) = -RL(T, c, k - 1);
            MTX(M, c, k, 0) = -(RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime) + RL(T, c, k - 1);
            MTX(M, c, k, 1) = 0.0;
        }
    }
    /* Temp matrix :237-261 */
    {
        int k = 1;
        for (size_t c = 0; c < NC; c++) {
            MTX(TM, c, k, -1) = 0.0;
            MTX(TM, c, k, 0) = -CpDry * (RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime)
                               + CpDry * RL(rExner, c, k) / ZL(zExner, c, k) * RL(TempTC, c, k);
            MTX(TM, c, k, 1) = -CpDry * RL(rExner, c, k) / ZL(zExner, c, k + 1) * RL(TempTC, c, k);
        }
        for (k = 2; k <= K - 1; k++)
            for (size_t c = 0; c < NC; c++) {
                MTX(TM, c, k, -1) = -CpDry * RL(rExner, c, k - 1) / ZL(zExner, c, k - 1) * RL(TempTC, c, k - 1);
                MTX(TM, c, k, 0) = -CpDry * (RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime)
                                   + CpDry * RL(rExner, c, k - 1) / ZL(zExner, c, k) * RL(TempTC, c, k - 1)
                                   + CpDry * RL(rExner, c, k) / ZL(zExner, c, k) * RL(TempTC, c, k);
                MTX(TM, c, k, 1) = -CpDry * RL(rExner, c, k) / ZL(zExner, c, k + 1) * RL(TempTC, c, k);
            }
        k = K;
        for (size_t c = 0; c < NC; c++) {
            MTX(TM, c, k, -1) = -CpDry * RL(rExner, c, k - 1) / ZL(zExner, c, k - 1) * RL(TempTC, c, k - 1);
            MTX(TM, c, k, 0) = -CpDry * (RL(Press, c, k) - RL(Press, c, k - 1)) / Grav / (2.0 * DelTime)
                               + CpDry * RL(rExner, c, k - 1) / ZL(zExner, c, k) * RL(TempTC, c, k - 1);
            MTX(TM, c, k, 1) = 0.0;
        }
    }

    /* RHS = - flux divergence :297-311 (the reference's OpenMP regions) */
    #pragma omp parallel
    {
        #pragma omp for
        for (int k = 1; k <= K; k++)
            for (size_t c = 0; c < NC; c++) {
                ZL(DUDt, c, k) = -(RL(MomFluxX, c, k) - RL(MomFluxX, c, k - 1));
                ZL(DVDt, c, k) = -(RL(MomFluxY, c, k) - RL(MomFluxY, c, k - 1));
                ZL(DTempDt, c, k) = -(RL(HeatFlux, c, k) - RL(HeatFlux, c, k - 1));
            }
        #pragma omp for collapse(2)
        for (int n = 0; n < ncmax; n++)
            for (int k = 1; k <= K; k++) {
                const double *F = QMixFlux + NC * (size_t)(K + 1) * n;
                double *D = DQMixDt + NC * (size_t)K * n;
                for (size_t c = 0; c < NC; c++) ZL(D, c, k) = -(RL(F, c, k) - RL(F, c, k - 1));
            }
    }

    for (size_t c = 0; c < NC; c++) {                                             /* :313-316 */
        Coef2[c + NC * 0] = ZL(DUDt, c, 1);
        Coef2[c + NC * 1] = ZL(DVDt, c, 1);
        Coef2[c + NC * 2] = ZL(DTempDt, c, 1);
        Coef2[c + NC * 3] = ZL((DQMixDt + NC * (size_t)K * iq), c, 1);
    }

    memcpy(Save, UV, sizeof(double) * NC * K * 3);                                /* :325 */
    tridiag_forward(NC, K, UV, DUDt);                                             /* :326 */
    tridiag_forward(NC, K, Save, DVDt);                                           /* :327 */
    tridiag_forward(NC, K, TM, DTempDt);                                          /* :328 */
    memcpy(Save, QM, sizeof(double) * NC * K * 3);                                /* :330 */
    for (int n = 0; n < ncmax; n++) {                                             /* :331-334 */
        memcpy(QM, Save, sizeof(double) * NC * K * 3);
        tridiag_forward(NC, K, QM, DQMixDt + NC * (size_t)K * n);
    }

    for (size_t c = 0; c < NC; c++) {                                             /* :344-376 */
        double tmp = -(RL(Press, c, 1) - RL(Press, c, 0)) / Grav / (2.0 * DelTime);
        double DFADUV1 = RL(VelTC, c, 1), DFADUV2 = -RL(VelTC, c, 1);
        Coef1[c + NC * 0] = tmp + DFADUV1 - DFADUV2 / MTX(UV, c, 2, 0);
        Coef2[c + NC * 0] = Coef2[c + NC * 0] - DFADUV2 * ZL(DUDt, c, 2) / MTX(UV, c, 2, 0);
        Coef1[c + NC * 1] = tmp + DFADUV1 - DFADUV2 / MTX(UV, c, 2, 0);
        Coef2[c + NC * 1] = Coef2[c + NC * 1] - DFADUV2 * ZL(DVDt, c, 2) / MTX(UV, c, 2, 0);
        double DFADT1 = CpDry * RL(rExner, c, 1) * RL(TempTC, c, 1) / ZL(zExner, c, 1);
        double DFADT2 = -CpDry * RL(rExner, c, 1) * RL(TempTC, c, 1) / ZL(zExner, c, 2);
        Coef1[c + NC * 2] = CpDry * tmp + DFADT1 - DFADT2 / MTX(TM, c, 2, 0);
        Coef2[c + NC * 2] = Coef2[c + NC * 2] - DFADT2 * ZL(DTempDt, c, 2) / MTX(TM, c, 2, 0);
        double DFADQ1 = RL(QMixTC, c, 1), DFADQ2 = -RL(QMixTC, c, 1);
        const double *DQ = DQMixDt + NC * (size_t)K * iq;
        Coef1[c + NC * 3] = tmp + DFADQ1 - DFADQ2 / MTX(QM, c, 2, 0);
        Coef2[c + NC * 3] = Coef2[c + NC * 3] - DFADQ2 * ZL(DQ, c, 2) / MTX(QM, c, 2, 0);
    }
    free(VelTC); free(TempTC); free(QMixTC); free(Tmp); free(Save);
}

/* ref :25-70 (SfcImplicitCoupling_VDiffBackward) */
void orc_vdiff_backward(orc_vdiff *h, double *DUDt, double *DVDt, double *DTempDt, double *DQMixDt)
{
    const size_t NC = (size_t)h->imax * h->jmax;
    const int K = h->kmax;
    const double DelTime = h->DelTime;
    tridiag_backward(NC, K, h->UVMtx, DUDt);                                      /* :51-53 */
    tridiag_backward(NC, K, h->UVMtx, DVDt);
    tridiag_backward(NC, K, h->TempMtx, DTempDt);
    for (int k = 1; k <= K; k++)                                                  /* :54-58 */
        for (size_t c = 0; c < NC; c++) {
            ZL(DUDt, c, k) = ZL(DUDt, c, k) / (2.0 * DelTime);
            ZL(DVDt, c, k) = ZL(DVDt, c, k) / (2.0 * DelTime);
            ZL(DTempDt, c, k) = ZL(DTempDt, c, k) / (2.0 * DelTime);
        }
    for (int n = 0; n < h->ncmax; n++) {                                          /* :60-63 */
        double *D = DQMixDt + NC * (size_t)K * n;
        tridiag_backward(NC, K, h->QMixMtx, D);
        for (size_t k = 0; k < NC * (size_t)K; k++) D[k] = D[k] / (2.0 * DelTime);
    }
}

void orc_vdiff_get_diag(const orc_vdiff *h, int which, double *out)
{
    const size_t NC = (size_t)h->imax * h->jmax;
    const int K = h->kmax;
    const double *M = which == 0 ? h->UVMtx : which == 1 ? h->TempMtx : h->QMixMtx;
    memcpy(out, M + NC * (size_t)K * 1, sizeof(double) * NC * K);
}

/* ------------------------------------------------------------------ ocean / sea-ice glue */

/* ref ocn/dccm_ocn_mod.f90:825-836: ice-surface selection by IceMaskMin (degC2K, IceMaskMin: DSIce) */
void orc_ocn_put_assemble(int64_t n, const double *SeaSfcTemp, const double *AlbAO, const double *SIceCon,
                          const double *SIceSfcTempC, const double *AlbAI, double IceMaskMin, double degC2K,
                          double *SIceSfcTemp, double *SIceAlbedo)
{
    for (int64_t c = 0; c < n; c++) {
        if (SIceCon[c] >= IceMaskMin) {
            SIceSfcTemp[c] = SIceSfcTempC[c] + degC2K;
            SIceAlbedo[c] = AlbAI[c];
        } else {
            SIceSfcTemp[c] = SeaSfcTemp[c];
            SIceAlbedo[c] = AlbAO[c];
        }
    }
}

/* ref ocn/dccm_ocn_mod.f90:978-993 */
void orc_ocn_get_assemble(int64_t n, const double *ns, const double *sr, const double *dFdT,
                          const double *Snow, const double *Rain, const double *EvapAO,
                          const double *WSXAO, const double *WSYAO, double DensFreshWater,
                          double *FreshWtFlxS0, double *FreshWtFlx0, double *WSXAI, double *WSYAI,
                          double *SfcHFlxAO0, double *DSfcHFlxAODTs)
{
    for (int64_t c = 0; c < n; c++) {
        FreshWtFlxS0[c] = ((Rain[c] + Snow[c]) - EvapAO[c]) / DensFreshWater;
        FreshWtFlx0[c] = FreshWtFlxS0[c];
        WSXAI[c] = WSXAO[c];
        WSYAI[c] = WSYAO[c];
        SfcHFlxAO0[c] = ns[c] + sr[c];
        DSfcHFlxAODTs[c] = dFdT[c];
    }
}

/* Jcup RECV_MODE='AVG' restated (ref ocn/dccm_ocn_mod.f90:652-672 registers every S->O / S->I variable with it;
 * Jcup itself is not part of the reference tree): running sum over the puts of an interval, divided by their number. */
void orc_avg_accumulate(double *acc, const double *x, int64_t n, int first)
{
    for (int64_t c = 0; c < n; c++) acc[c] = first ? x[c] : acc[c] + x[c];
}

void orc_avg_finish(double *acc, int64_t n, int count)
{
    for (int64_t c = 0; c < n; c++) acc[c] = acc[c] / (double)count;
}

/* dcpam_StoreAtmSurfFlxInfo, ref atm/dcpam_main_mod.f90:1068-1112 (array syntax -> one loop; the saturation
 * derivatives of :1063-1067 come from DCPAM's saturate module and are inputs).  in[] / out[] follow the member
 * order of dccm_atm_sfcflx in include/dccm_b200.h. */
void orc_atm_store_surf_flx(int64_t n, const double *const *in, double *const *out,
                            double LatentHeat, double CpDry, double delta_t)
{
    const double *SurfMomFluxX = in[0], *SurfMomFluxY = in[1], *VelTC = in[2], *TempTC = in[3], *QVapTC = in[4],
                 *HumidCoef = in[5], *DUDt1 = in[6], *DVDt1 = in[7], *DTempDt1 = in[8], *DQVapDt1 = in[9],
                 *HeatFlux0 = in[10], *QVapFlux0 = in[11], *ExnerR0 = in[12], *ExnerZ1 = in[13], *TempN1 = in[14],
                 *DSurfTempDt = in[15], *SnowFrac = in[16], *DQOnLiq = in[17], *DQOnSol = in[18],
                 *RadLDw0 = in[19], *RadLUw0 = in[20], *RadSDw0 = in[21], *RadSUw0 = in[22],
                 *DelLDw00 = in[23], *DelLDw01 = in[24], *DelLUw00 = in[25], *DelLUw01 = in[26];
    for (int64_t c = 0; c < n; c++) {
        const double dqsat = (1.0 - SnowFrac[c]) * DQOnLiq[c] + SnowFrac[c] * DQOnSol[c];
        out[0][c] = SurfMomFluxX[c] - VelTC[c] * DUDt1[c] * 2.0 * delta_t;
        out[1][c] = SurfMomFluxY[c] - VelTC[c] * DVDt1[c] * 2.0 * delta_t;
        out[2][c] = HeatFlux0[c] - CpDry * ExnerR0[c] * TempTC[c]
                    * (DTempDt1[c] / ExnerZ1[c] - DSurfTempDt[c] / ExnerR0[c]) * 2.0 * delta_t;
        out[3][c] = LatentHeat * (QVapFlux0[c]
                    - HumidCoef[c] * QVapTC[c] * (DQVapDt1[c] - dqsat * DSurfTempDt[c]) * 2.0 * delta_t);
        out[4][c] = RadLDw0[c] + 2.0 * delta_t * (DSurfTempDt[c] * DelLDw00[c] + DTempDt1[c] * DelLDw01[c]);
        out[5][c] = RadLUw0[c] + 2.0 * delta_t * (DSurfTempDt[c] * DelLUw00[c] + DTempDt1[c] * DelLUw01[c]);
        out[6][c] = RadSDw0[c];
        out[7][c] = RadSUw0[c];
        out[8][c] = ExnerR0[c] / ExnerZ1[c] * TempN1[c];
        out[9][c] = LatentHeat * HumidCoef[c] * QVapTC[c] * dqsat;
        out[10][c] = CpDry * TempTC[c] + out[9][c] - DelLDw00[c];
    }
}

/* ref atm/dccm_atm_mod.f90:831: xy_SfcTemp(:,:) = (xy_LUwRFlx/StB)**0.25d0 -- the radiative surface temperature the
 * atmosphere derives from the remapped composite upward long-wave flux */
void orc_atm_sfc_temp(int64_t n, const double *LUwRFlx, double StB, double *SfcTemp)
{
    for (int64_t c = 0; c < n; c++) SfcTemp[c] = m_pow(LUwRFlx[c] / StB, 0.25);
}

/* legacy 2-component get side: ref atm/mod_atm.f90:743 (fourth root), :772-773 (snow x 1e3, flux residual x coupling
 * cycle) and dcpam_UpdateSurfaceProperties, ref atm/dcpam_main_mod.f90:1026-1028 (lowest-level temperature correction) */
void orc_atm_legacy_get(int64_t n, const double *SfcTemp4, const double *SfcSnow, const double *SfcEngyFlxMod,
                        double cycle_sec, double Grav, double CpDry, const double *Press0, const double *Press1,
                        double *SurfTemp, double *SurfSnow, double *TempB1)
{
    for (int64_t c = 0; c < n; c++) {
        SurfTemp[c] = m_pow(SfcTemp4[c], 0.25);
        SurfSnow[c] = 1e3 * SfcSnow[c];
        double recv = SfcEngyFlxMod[c] * cycle_sec;
        TempB1[c] = TempB1[c] + (recv - 0.0) / (Press0[c] - Press1[c]) * Grav / CpDry;
    }
}

