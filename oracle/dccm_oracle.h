/*
 * dccm_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the Dennou-CCM surface-exchange hot path, following the
 * reference Fortran loop nests line by line (citations are relative to
 * /root/reference and given on every function in dccm_oracle.c).
 *
 * PARITY UNPINNED: the reference ships no tests, no golden vectors and no mapping
 * tables, and it is Fortran-only (no Fortran compiler in this image; the hot-path
 * modules also `use` the un-vendored gtool5 / Jcup / DCPAM modules), so neither
 * golden fixtures nor an `oracle/_ref` build of the real reference exist.  The
 * oracle is pinned instead by self-derived known-answer tests (tests/test_oracle_kat.py,
 * SURVEY.md section 4 KAT-1..5).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 */
#ifndef DCCM_ORACLE_H
#define DCCM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- mapping-table container (one entry = one text line "iD jD iS jS coef") ---- */
typedef struct orc_table {
    int64_t n, cap;
    int32_t *iD, *jD, *iS, *jS;   /* 1-based, as written to the table file */
    double  *coef;
} orc_table;

orc_table *orc_table_new(void);
void       orc_table_free(orc_table *t);
int64_t    orc_table_n(const orc_table *t);
void       orc_table_copy(const orc_table *t, int32_t *iD, int32_t *jD, int32_t *iS,
                          int32_t *jS, double *coef);

/* ---- grids (stand-in for SPML w_module) ---- */
int orc_gauss_legendre(int n, double *mu, double *w);
int orc_gauss_grid(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt);
int orc_regular_grid(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt);

/* ---- table generators ---- */
/* lon_mode 0 = reference-faithful; 1 = generalised longitude overlap (extension, no parity claim) */
int orc_gen_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                    int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                    const double *y_LatIntWtS, const double *y_LatIntWtD,
                    int accuracy_order, int lon_mode, orc_table *out);
int orc_gen_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                     int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                     int lon_mode, orc_table *out);
/* the same generators restricted to destination rows j0..j1 (1-based, inclusive): exactly those lines of the full table */
int orc_gen_jones99_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                         int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                         const double *y_LatIntWtS, const double *y_LatIntWtD,
                         int accuracy_order, int lon_mode, int jD0, int jD1, orc_table *out);
int orc_gen_bilinear_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                          int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                          int lon_mode, int jr0, int jr1, orc_table *out);
/* make_mapping_table of the stand-alone regular-grid generator, ref common/cal_mappingtable.f90:10-49 */
int orc_make_mapping_table(int nx_r, int ny_r, int nx_s, int ny_s, orc_table *out);
int orc_exchange_grid(int jma, const double *y_LatA, const double *y_IntWtLatA,
                      int jmo, const double *y_IntWtLatO,
                      int *jms, double *y_LatS, double *y_IntWtLatS);

/* ---- table file I/O ---- */
int orc_table_write_text(const orc_table *t, const char *filename);
int orc_table_read_text(const char *filename, orc_table *out);
void orc_table_to_index(const orc_table *t, int gnxs, int gnxr,
                        int32_t *send_index, int32_t *recv_index, double *coef_s);

/* ---- remap apply ---- */
void orc_remap_apply(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                     const double *coef, const double *send, int sn1, int sn2,
                     double *recv, int rn1, int rn2, int num_of_data);

/* ---- elementary functions of the bulk flux (orc_pmath.h): 1 = portable fixed IEEE sequences (default), 0 = libm ---- */
void orc_set_math(int portable);
int  orc_get_math(void);
void orc_pm_exp_v(int64_t n, const double *x, double *y);
void orc_pm_log_v(int64_t n, const double *x, double *y);
void orc_pm_pow_v(int64_t n, const double *x, double e, double *y);

/* ---- bulk flux ---- */
void orc_bulkflux(int IA, int JA,
    double *xya_WindStressX, double *xya_WindStressY,
    double *xya_SenHFlx, double *xya_QVapMFlx, double *xya_LatHFlx,
    double *xya_SfcVelTransCoef, double *xya_SfcTempTransCoef, double *xya_SfcQVapTransCoef,
    double *xya_DelVarImplCPL,
    double *xya_SUwRFlx, double *xya_LUwRFlx,
    double *xya_SfcHFlx_ns, double *xya_SfcHFlx_sr, double *xya_DSfcHFlxDTs,
    const double *xy_WindU, const double *xy_WindV, const double *xy_SfcAirTemp, const double *xy_QVap1,
    const double *xy_SDwRFlx, const double *xy_LDwRFlx,
    const double *xya_ImplCplCoef1, const double *xya_ImplCplCoef2,
    double *xya_SfcTemp, double *xya_SfcAlbedo, const double *xy_SIceCon,
    const double *a_Sig1Info, const double *xy_SfcHeight, const double *xy_SfcPress);

/* ---- implicit coupling (column tridiagonal) ---- */
typedef struct orc_vdiff {
    int imax, jmax, kmax, ncmax, index_h2ovap; /* index_h2ovap 1-based */
    double Grav, CpDry, GasRDry, DelTime;
    double *UVMtx, *TempMtx, *QMixMtx;         /* (imax*jmax, kmax, 3) */
} orc_vdiff;

orc_vdiff *orc_vdiff_new(int imax, int jmax, int kmax, int ncmax, int index_h2ovap,
                         double Grav, double CpDry, double GasRDry, double DelTime);
void orc_vdiff_free(orc_vdiff *h);
void orc_vdiff_forward(orc_vdiff *h,
    const double *xyr_MomFluxX, const double *xyr_MomFluxY, const double *xyr_HeatFlux,
    const double *xyrf_QMixFlux,
    const double *xyr_Press, const double *xyz_Exner, const double *xyr_Exner,
    const double *xyr_VirTemp, const double *xyz_Height,
    const double *xyr_VelDiffCoef, const double *xyr_TempDiffCoef, const double *xyr_QMixDiffCoef,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    double *xya_ImplCplCoef1, double *xya_ImplCplCoef2);
void orc_vdiff_backward(orc_vdiff *h,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt);
/* copy of the swept diagonal b'(k) (k = 1..kmax; k=1 unspecified, see defect C-1) */
void orc_vdiff_get_diag(const orc_vdiff *h, int which, double *out);

/* ---- ocean / sea-ice glue, element-wise (ref ocn/dccm_ocn_mod.f90:825-836, :978-993) ---- */
void orc_ocn_put_assemble(int64_t n, const double *SeaSfcTemp, const double *AlbAO, const double *SIceCon,
                          const double *SIceSfcTempC, const double *AlbAI, double IceMaskMin, double degC2K,
                          double *SIceSfcTemp, double *SIceAlbedo);
void orc_ocn_get_assemble(int64_t n, const double *ns, const double *sr, const double *dFdT,
                          const double *Snow, const double *Rain, const double *EvapAO,
                          const double *WSXAO, const double *WSYAO, double DensFreshWater,
                          double *FreshWtFlxS0, double *FreshWtFlx0, double *WSXAI, double *WSYAI,
                          double *SfcHFlxAO0, double *DSfcHFlxAODTs);

/* ---- Jcup RECV_MODE='AVG' and dcpam_StoreAtmSurfFlxInfo (ref atm/dcpam_main_mod.f90:1068-1112) ---- */
void orc_avg_accumulate(double *acc, const double *x, int64_t n, int first);
void orc_avg_finish(double *acc, int64_t n, int count);
void orc_atm_store_surf_flx(int64_t n, const double *const *in, double *const *out,
                            double LatentHeat, double CpDry, double delta_t);

/* ref atm/dccm_atm_mod.f90:831 */
void orc_atm_sfc_temp(int64_t n, const double *LUwRFlx, double StB, double *SfcTemp);

/* ref atm/mod_atm.f90:743, :772-773; atm/dcpam_main_mod.f90:1026-1028 */
void orc_atm_legacy_get(int64_t n, const double *SfcTemp4, const double *SfcSnow, const double *SfcEngyFlxMod,
                        double cycle_sec, double Grav, double CpDry, const double *Press0, const double *Press1,
                        double *SurfTemp, double *SurfSnow, double *TempB1);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
