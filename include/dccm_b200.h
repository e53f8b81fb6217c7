/*
 * dccm_b200.h -- C ABI of the B200-native surface-exchange path of Dennou-CCM.
 *
 * This is the drop-in boundary: every entry point is what an ISO_C_BINDING shim with the
 * reference's UNCHANGED Fortran interface binds to (shims: fortran/, wiring: INTEGRATION.md).
 * Plain pointers and sizes only; default Fortran INTEGER = int32_t, REAL(DP) = double;
 * all arrays are Fortran column-major; indices stored in tables are 1-based as in the
 * reference.  `ref:` citations are file:line under the reference tree.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; dccm_last_error() gives the
 *     message (thread-local).  There is NO CPU fallback: a call that needs the GPU fails
 *     with DCCM_ERR_CUDA when no sm_100 device is usable.
 *   - *_host entry points take HOST buffers and are synchronous (results are in the host
 *     arrays on return), i.e. exactly the reference semantics.
 *   - *_device entry points take DEVICE pointers and a cudaStream_t (as void*) and are
 *     asynchronous on that stream; they are what the resident exchange step is built from.
 */
#ifndef DCCM_B200_H
#define DCCM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCCM_OK            0
#define DCCM_ERR_ARG       1
#define DCCM_ERR_CUDA      2
#define DCCM_ERR_IO        3
#define DCCM_ERR_UNSUPPORTED 4   /* grid pair the reference generator cannot handle (lon_mode 0) */
#define DCCM_ERR_SEARCH    5     /* "Exception.." stop of search_OverwrapRange */

/* ------------------------------------------------------------------ runtime */
const char *dccm_last_error(void);
const char *dccm_build_info(void);
int dccm_init(int device);                 /* cudaSetDevice + capability check (sm_100) */
int dccm_device_count(int *count);
int dccm_sync(void *stream);               /* cudaStreamSynchronize */

/* Page-lock / release a caller-owned host array (e.g. the module arrays of DSFCM_Admin_Variable_mod or DCPAM's
 * tendency arrays, once at init).  The *_host entry points move their arguments in chunks with H2D, kernel and D2H
 * overlapped on three streams; that overlap needs page-locked memory (pageable arrays still work, serialised). */
int dccm_host_register(void *ptr, int64_t bytes);
int dccm_host_unregister(void *ptr);

/* ------------------------------------------------------------------ grids
 * Stand-ins for the SPML w_module axes gmapgen uses (ref tool/gmapgen/gmapgen_main.f90:256-307). */
int dccm_grid_gauss(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt);
int dccm_grid_regular(int im, int jm, double *x_Lon, double *y_Lat, double *x_LonWt, double *y_LatWt);
/* generate_surface_exchage_grid, ref tool/gmapgen/gmapgen_main.f90:336-405.
 * y_LatS / y_IntWtLatS need room for jma+jmo-1 values; longitudes are the atmosphere's. */
int dccm_grid_exchange(int jma, const double *y_LatA, const double *y_IntWtLatA,
                       int jmo, const double *y_IntWtLatO,
                       int *jms, double *y_LatS, double *y_IntWtLatS);

/* ------------------------------------------------------------------ mapping tables */
typedef struct dccm_table dccm_table;

/* gen_gridmapfile_lonlat2lonlat (Jones 1999), ref common/grid_mapping_util_jones99.f90:35-442.
 * lon_mode 0: reference behaviour (equal longitudes or nx==1 on one side; anything else ->
 * DCCM_ERR_UNSUPPORTED).  lon_mode 1: generalised longitude overlap for mismatched longitudes
 * (first order only); identical to mode 0 on the grid pairs the reference supports. */
int dccm_table_gen_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                           int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                           const double *y_LatIntWtS, const double *y_LatIntWtD,
                           int accuracy_order, int lon_mode, dccm_table **out);
/* gen_gridmapfile_lonlat2lonlat (bilinear), ref common/grid_mapping_util.f90:32-177.
 * lon_mode 1 unwraps the east neighbour by 2*pi at the wrap column (ref :117-119 does not). */
int dccm_table_gen_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                            int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                            int lon_mode, dccm_table **out);
/* make_mapping_table of the older stand-alone bilinear generator, ref common/cal_mappingtable.f90:10-49 (cal_coef
 * :59-76): REGULAR grids given by their sizes (degrees: longitudes 0, 360/nx, ...; latitudes pole to pole in steps of
 * 180/(ny-1)), receiver (nx_r, ny_r) <- sender (nx_s, ny_s).  Entries with coefficient > 0 only, in the order
 * (is,js) (is+1,js) (is+1,js+1) (is,js+1) with mod wrap, as the reference writes them (:40-43).  Host only. */
int dccm_table_gen_make_mapping_table(int nx_r, int ny_r, int nx_s, int ny_s, dccm_table **out);
/* Same generators restricted to destination rows jd_first..jd_last (1-based, inclusive): the
 * entries a rank owning that latitude band needs (row-block sharding).  Entries are
 * identical to the corresponding slice of the full table. */
int dccm_table_gen_jones99_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                const double *y_LatIntWtS, const double *y_LatIntWtD,
                                int accuracy_order, int lon_mode, int jd_first, int jd_last,
                                dccm_table **out);
int dccm_table_gen_bilinear_rows(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                 int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                 int lon_mode, int jr_first, int jr_last, dccm_table **out);
/* The same tables produced by multiplying out the separable factors (longitude list per destination column x
 * latitude list per destination row) the way the kind-2 kernels do: identical to the generators' output entry for
 * entry wherever the separable form applies (Jones99: different longitudes, 1st order; bilinear: nx > 1 on both
 * sides), DCCM_ERR_UNSUPPORTED otherwise.  Host only. */
int dccm_table_gen_jones99_separable(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                     int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                     const double *y_LatIntWtS, const double *y_LatIntWtD,
                                     int accuracy_order, int lon_mode, dccm_table **out);
int dccm_table_gen_bilinear_separable(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                      int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                      int lon_mode, dccm_table **out);
/* The reference's on-disk format: one entry per line "iD jD iS jS coef", list-directed
 * (ref common/grid_mapping_util_jones99.f90:246-247, :479-504). */
int dccm_table_write_text(const dccm_table *t, const char *filename);
int dccm_table_read_text(const char *filename, dccm_table **out);
/* compact binary form of the same entries */
int dccm_table_write_bin(const dccm_table *t, const char *filename);
int dccm_table_read_bin(const char *filename, dccm_table **out);
int64_t dccm_table_size(const dccm_table *t);
int dccm_table_get(const dccm_table *t, int32_t *iD, int32_t *jD, int32_t *iS, int32_t *jS, double *coef);
/* set_mappingTable_interpCoef, ref common/grid_mapping_util_jones99.f90:446-506
 * (recv_index = iD + GNXR*(jD-1), send_index = iS + GNXS*(jS-1)). */
int dccm_table_index(const dccm_table *t, int gnxs, int gnxr,
                     int32_t *send_index, int32_t *recv_index, double *coef_s);
void dccm_table_free(dccm_table *t);

/* ------------------------------------------------------------------ remap apply (K1)
 * Replaces interpolate_data / interpolate_data_latlon,
 * ref common/interpolate_data.f90:1-17, common/interpolation_data_latlon_mod.f90:274-306. */
typedef struct dccm_remap dccm_remap;

/* Built once where the reference calls set_operation_index + set_interpolate_coef
 * (ref common/interpolation_data_latlon_mod.f90:116-154, :219-270): local 1-based
 * send/recv indices and coefS in operation (= table) order. */
int dccm_remap_create(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                      const double *coef, int n_send, int n_recv, dccm_remap **out);
/* Same, with the i-fastest row lengths of the two grids (GNXS, GNXR of set_mappingTable_interpCoef,
 * ref common/grid_mapping_util_jones99.f90:446-462).  When every destination latitude row applies one
 * longitude-shifted stencil -- true for all tables the reference generator produces -- the table is
 * stored as that O(ny) stencil (dccm_remap_kind() == 1) instead of O(nx*ny) CSR.  Results are
 * bit-identical either way; gnxs = gnxr = 0 skips the detection. */
int dccm_remap_create_lonlat(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                             const double *coef, int n_send, int n_recv, int gnxs, int gnxr,
                             dccm_remap **out);
/* Generate-and-create in one step: same arguments as the generators above, the operator comes
 * back without the table ever being built.  Grid pairs with different longitudes (lon_mode 1) give a SEPARABLE
 * operator (dccm_remap_kind() == 2): per-column longitude factors and per-row latitude factors, O(nx + ny) numbers
 * that the kernels multiply out exactly as the generator would have (same order, same product, same 1e-14 drop
 * test) -- bit-identical results, no O(nx*ny) table in memory or in the kernels' traffic.  Equal longitudes and an
 * axisymmetric source (any accuracy order) come back directly as one stencil per destination row (kind 1), again without
 * a table; what is left (axisymmetric destination, very long stencils) goes through the generator and
 * dccm_remap_create_lonlat. */
int dccm_remap_create_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                              int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                              const double *y_LatIntWtS, const double *y_LatIntWtD,
                              int accuracy_order, int lon_mode, dccm_remap **out);
int dccm_remap_create_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                               int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                               int lon_mode, dccm_remap **out);
/* The same operators for a LATITUDE BAND of the destination grid (row-block sharding over GPUs; the
 * reference partitions the remap by the receiving component's local cells, ref common/interpolation_data_latlon_mod.f90:
 * 140-151): destination rows [jD0, jD1) (0-based, half open) become rows 0.. of the operator, and source cells are
 * numbered inside the band's source buffer, which holds the src_rows source rows starting at row src_row0 (own rows +
 * halo).  Exactly the lines of the full table that belong to those rows, in the same storage forms (zonal stencils /
 * separable factors, no table built) -- results are bit-identical to the corresponding rows of the full operator. */
int dccm_remap_create_jones99_band(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                   int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                   const double *y_LatIntWtS, const double *y_LatIntWtD,
                                   int accuracy_order, int lon_mode,
                                   int jD0, int jD1, int src_row0, int src_rows, dccm_remap **out);
int dccm_remap_create_bilinear_band(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                    int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                    int lon_mode, int jD0, int jD1, int src_row0, int src_rows, dccm_remap **out);
/* host-only: what a band operator applies, multiplied out into a table with the band's LOCAL indices (checks) */
int dccm_table_gen_band_expanded(int conservative, int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                 int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                 const double *y_LatIntWtS, const double *y_LatIntWtD,
                                 int accuracy_order, int lon_mode,
                                 int jD0, int jD1, int src_row0, int src_rows, dccm_table **out);
/* host-only: the storage form dccm_remap_create_lonlat would choose (kind, entries kept) */
int dccm_remap_classify(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                        const double *coef, int n_send, int n_recv, int gnxs, int gnxr,
                        int *kind, int64_t *stored_entries);
void dccm_remap_destroy(dccm_remap *h);
int64_t dccm_remap_nnz(const dccm_remap *h);
/* 0 = general destination-row CSR, 1 = zonal stencil (one stencil per destination latitude row),
 * 2 = separable (longitude factors x latitude factors, dccm_remap_create_jones99 / _bilinear) */
int dccm_remap_kind(const dccm_remap *h);
/* recv_data(:,:) = 0 ; recv(r_i,d) += send(s_i,d)*coef(i), d = 1..num_of_data  (ref :293-302) */
int dccm_remap_apply_host(dccm_remap *h, const double *send, int sn1, int sn2,
                          double *recv, int rn1, int rn2, int num_of_data);
int dccm_remap_apply_device(dccm_remap *h, const double *d_send, int sn1,
                            double *d_recv, int rn1, int rn2, int num_of_data, void *stream);
/* Sharded runs: a source buffer whose first b0 cells are the lower neighbour's boundary rows and whose
 * cells from b1 on are the upper neighbour's, read IN PLACE from the neighbours' own buffers through
 * NVLink peer mappings (CUDA IPC / symmetric memory) -- the remap is fused with its halo "collective".
 * lo / own / hi are pre-offset so that element (cell c, layer l) is base[c + l*sn1]. */
typedef struct dccm_src_seg {
    const double *lo, *own, *hi;
    int64_t b0, b1;
} dccm_src_seg;
int dccm_remap_apply_seg_device(dccm_remap *h, const dccm_src_seg *send, int sn1,
                                double *d_recv, int rn1, int rn2, int num_of_data, void *stream);
/* registry keyed like the reference's operation_index(recv_model, send_model, mapping_tag)
 * (ref :88, :289); ids are Jcup component numbers (1-based), tags as in
 * common/dccm_common_params_mod.f90:64-69. */
int dccm_interp_register(int recv_model, int send_model, int mapping_tag, dccm_remap *h);
int dccm_interpolate_data(int recv_model, int send_model, int mapping_tag,
                          int sn1, int sn2, const double *send_data,
                          int rn1, int rn2, double *recv_data, int num_of_data);

/* The coupler's callback WITHOUT a Fortran shim.  Jcup resolves the bare external symbol `interpolate_data_`
 * (ref common/interpolate_data.f90:1-17; linked as a lone object, Mkinclude:167).  The library exports that symbol
 * itself with the Fortran calling convention of gfortran / ifort / nvfortran on 64-bit Linux -- every argument by
 * reference, the two CHARACTER(*) lengths appended by value -- so leaving interpolate_data.o out of the link line
 * and adding -ldccm_b200 is enough.  Component names are translated through a table the host fills once with
 * what jcup_get_comp_num_from_name returns (ref common/interpolation_data_latlon_mod.f90:289).  The subroutine has
 * no status argument: a failure goes to the error handler (default: message to stderr and abort(), the behaviour of
 * the reference's jcup_error); dccm_interpolate_data_named is the same call with a status for C callers. */
int dccm_interp_set_model_name(int model_id, const char *name);
int dccm_interpolate_data_named(const char *recv_model, int64_t recv_len,
                                const char *send_model, int64_t send_len, int mapping_tag,
                                int sn1, int sn2, const double *send_data,
                                int rn1, int rn2, double *recv_data, int num_of_data);
void dccm_f77_set_error_handler(void (*handler)(const char *message));
void interpolate_data_(const char *recv_model, const char *send_model, const int32_t *mapping_tag,
                       const int32_t *sn1, const int32_t *sn2, const double *send_data,
                       const int32_t *rn1, const int32_t *rn2, double *recv_data,
                       const int32_t *num_of_data, const int32_t *tn, const int32_t *exchange_tag,
                       size_t recv_model_len, size_t send_model_len);

/* ------------------------------------------------------------------ bulk flux (K2)
 * Replaces DSFCM_Util_SfcBulkFlux_Get, ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:108-439
 * (+ BulkCoefL82 :443-572, limits :574-591).  Argument order = the Fortran dummy order. */
int dccm_bulkflux_get_host(int IA, int JA,
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

/* Device form.  Columns are addressed as col(i,j) = off + i + ld*j, i<nx, j<ny; the n-th
 * slot of a 3-D array is at + n*slot_stride.  (IA,JA)-haloed arrays: nx=IA-2, ny=JA-2,
 * ld=IA, off=IA+1, slot_stride=IA*JA.  Output pointers may be NULL (not stored). */
typedef struct dccm_sfc_fields {
    /* out (3 slots each; DelVarImplCPL 4) */
    double *WindStressX, *WindStressY, *SenHFlx, *QVapMFlx, *LatHFlx;
    double *SfcVelTransCoef, *SfcTempTransCoef, *SfcQVapTransCoef;
    double *DelVarImplCPL;
    double *SUwRFlx, *LUwRFlx;
    double *SfcHFlx_ns, *SfcHFlx_sr, *DSfcHFlxDTs;
    /* in */
    const double *WindU, *WindV, *SfcAirTemp, *QVap1, *SDwRFlx, *LDwRFlx;
    const double *ImplCplCoef1, *ImplCplCoef2;   /* 4 slots */
    double *SfcTemp, *SfcAlbedo;                 /* inout: slots 1,2 in, slot 3 out */
    const double *SIceCon;
    const double *SfcHeight;                     /* NULL = 0 everywhere (sfc/dccm_sfc_mod.f90:885) */
    const double *SfcPress;
} dccm_sfc_fields;
int dccm_bulkflux_device(int nx, int ny, int ld, int64_t off, int64_t slot_stride,
                         const dccm_sfc_fields *f, double sig1, void *stream);

/* Self-test of the branch-free fp64 division / reciprocal / square root the bulk-flux kernels use
 * (csrc/dccm_bulkflux.cuh, FastArith): evaluates a[i]/b[i], 1/b[i], sqrt(|a[i]|) both ways on the device.
 * mismatches = accepted fast results whose bits differ from the plain operator (must be 0);
 * rejected = operations outside the fast paths' exponent range (re-evaluated with plain operators). */
int dccm_selftest_fast_arith_device(const double *d_a, const double *d_b, int64_t n,
                                    int64_t *mismatches, int64_t *rejected);

/* The elementary functions of the bulk flux (csrc/dccm_pmath.cuh: exp, log, x**y as fixed sequences of IEEE
 * binary64 operations -- DESIGN.md section 5), evaluated element-wise on device arrays so a host can check
 * its own text of the sequences against the device's bit for bit.
 * which: 0 = exp(x), 1 = log(x), 2 = x**y, 3 = x**0.25 (two square roots). */
int dccm_selftest_pmath_device(int which, const double *d_x, int64_t n, double y, double *d_out);

/* ------------------------------------------------------------------ fused surface step (K1+K2)
 * The surface component's whole coupling step, get -> bulk flux -> put, in one kernel:
 * replaces jcup_get_data x16 -> interpolate_data -> unpack (ref sfc/dccm_sfc_mod.f90:865-881,
 * :900-951), DSFCM_Util_SfcBulkFlux_Get (:886-897) and the put-side selection/pack (:764-809).
 * Buffers are layer-major with `members` ensemble members per layer (row = layer*members+m):
 *   a2s_bil (13 layers, nA): WindU WindV SfcAirTemp QVap1 SfcPress ImplCplCoef1(4) ImplCplCoef2(4)
 *   a2s_cons (4, nA): LDwRFlx SDwRFlx RainFall SnowFall
 *   o2s_bil (2, nO): SfcTemp(ocean) SfcTemp(ice)   o2s_cons (3, nO): SIceCon SfcAlbedo(ocean) SfcAlbedo(ice)
 *   s2a (9, nS): LUwRFlx SUwRFlx SenHFlx QVapMFlx [composite] | SfcAlbedo(3) DelVarImplCPL(4)
 *   s2o (12, nS): SfcHFlx_ns SfcHFlx_sr [ocean] SnowFall RainFall Evap -WindStressX -WindStressY |
 *                 SfcHFlx_ns SfcHFlx_sr Evap [ice] | DSfcHFlxDTs(ocean) DSfcHFlxDTs(ice)
 * s_ld = cells per row of s2a / s2o (0 = nS; larger when the rows also hold halo cells of a
 * sharded run, with s2a / s2o pointing at the first owned cell).
 * `full` (optional) receives the API-complete DSFCM arrays, slot stride members*nS. */
int dccm_sfc_exchange_device(const dccm_remap *as_bil, const dccm_remap *as_cons,
                             const dccm_remap *os_bil, const dccm_remap *os_cons,
                             const double *a2s_bil, const double *a2s_cons,
                             const double *o2s_bil, const double *o2s_cons,
                             int members, double sig1, double *s2a, double *s2o, int64_t s_ld,
                             const dccm_sfc_fields *full, void *stream);

/* Which form of the kernel the calls made with this A->S bilinear handle launch (the setting lives on the handle:
 * no process-wide state): staged = 1 (default) lets a CTA bring its atmosphere source rows to shared memory with
 * TMA bulk copies whenever both A->S tables are zonal stencils, 0 forces the per-thread global gathers;
 * min_blocks = CTAs per SM the kernel is compiled for (4, 5 or 6).  A negative value keeps the current setting.
 * Results are bit-identical in every configuration.  A handle (its redo list, these options) serves one stream
 * at a time: concurrent exchanges use separate handles. */
int dccm_sfc_exchange_config(dccm_remap *as_bil, int staged, int min_blocks);
/* 1 if the last dccm_sfc_exchange_*_device call with this handle launched the staged form, 0 the direct one, -1 none yet */
int dccm_sfc_exchange_last_form(const dccm_remap *as_bil);

/* same, source buffers given as peer-mapped segments; a_ld / o_ld = cells per row of the ATM / OCN
 * send buffers (0 = the tables' source extent) */
int dccm_sfc_exchange_seg_device(const dccm_remap *as_bil, const dccm_remap *as_cons,
                                 const dccm_remap *os_bil, const dccm_remap *os_cons,
                                 const dccm_src_seg *a2s_bil, const dccm_src_seg *a2s_cons,
                                 const dccm_src_seg *o2s_bil, const dccm_src_seg *o2s_cons,
                                 int64_t a_ld, int64_t o_ld,
                                 int members, double sig1, double *s2a, double *s2o, int64_t s_ld,
                                 const dccm_sfc_fields *full, void *stream);
/* same for rows [row0, row1) of the exchange grid only (row1 < 0: every row): lets a caller pipeline the exchange over
 * latitude slabs -- the surface kernel of slab k next to the column solve of slab k+2 on another stream
 * (exchange.SurfaceExchange.step_pipelined).  Needs structured A->S tables (row length known). */
int dccm_sfc_exchange_rows_device(const dccm_remap *as_bil, const dccm_remap *as_cons,
                                  const dccm_remap *os_bil, const dccm_remap *os_cons,
                                  const dccm_src_seg *a2s_bil, const dccm_src_seg *a2s_cons,
                                  const dccm_src_seg *o2s_bil, const dccm_src_seg *o2s_cons,
                                  int64_t a_ld, int64_t o_ld,
                                  int members, double sig1, double *s2a, double *s2o, int64_t s_ld,
                                  const dccm_sfc_fields *full, int row0, int row1, void *stream);

/* ------------------------------------------------------------------ implicit coupling (K3/K4)
 * Replaces dcpam_sfc_implicit_coupling_mod, ref atm/dcpam_sfc_implicit_coupling_mod.f90. */
typedef struct dccm_vdiff dccm_vdiff;

/* dcpam_sfc_implicit_coupling_Init (ref :420-426); sizes/constants come from DCPAM's
 * gridset / composition / constants / timeset modules (ref :3-7, :85-108).
 * index_h2ovap is 1-based.  The swept matrices live in the handle (ref :16-18 `save`). */
int dccm_vdiff_create(int imax, int jmax, int kmax, int ncmax, int index_h2ovap,
                      double Grav, double CpDry, double GasRDry, double DelTime, dccm_vdiff **out);
void dccm_vdiff_destroy(dccm_vdiff *h);
/* 0 = operation order of the reference (bit-exact vs the oracle), 1 = shared reciprocals */
int dccm_vdiff_set_mode(dccm_vdiff *h, int fast);
/* slot stride of xya_ImplCplCoef1/2 in the *_device form (0 = imax*jmax): lets the forward
 * solve write the coefficients straight into a wider A->S send buffer (sharded runs). */
int dccm_vdiff_set_coef_stride(dccm_vdiff *h, int64_t slot_stride);
/* SfcImplicitCoupling_VDiffForward (ref :72-378), Fortran dummy order. */
int dccm_vdiff_forward_host(dccm_vdiff *h,
    const double *xyr_MomFluxX, const double *xyr_MomFluxY, const double *xyr_HeatFlux,
    const double *xyrf_QMixFlux,
    const double *xyr_Press, const double *xyz_Exner, const double *xyr_Exner,
    const double *xyr_VirTemp, const double *xyz_Height,
    const double *xyr_VelDiffCoef, const double *xyr_TempDiffCoef, const double *xyr_QMixDiffCoef,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    double *xya_ImplCplCoef1, double *xya_ImplCplCoef2);
/* SfcImplicitCoupling_VDiffBackward (ref :25-70) */
int dccm_vdiff_backward_host(dccm_vdiff *h,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt);
int dccm_vdiff_forward_device(dccm_vdiff *h,
    const double *xyr_MomFluxX, const double *xyr_MomFluxY, const double *xyr_HeatFlux,
    const double *xyrf_QMixFlux,
    const double *xyr_Press, const double *xyz_Exner, const double *xyr_Exner,
    const double *xyr_VirTemp, const double *xyz_Height,
    const double *xyr_VelDiffCoef, const double *xyr_TempDiffCoef, const double *xyr_QMixDiffCoef,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    double *xya_ImplCplCoef1, double *xya_ImplCplCoef2, void *stream);
/* level1 (4 slots of imax*jmax: U,V,T,q_vap increments = DelVarImplCPL remapped to the ATM
 * grid) replaces level 1 of the tendencies first, as atm/dccm_atm_mod.f90:832-835 does; NULL
 * = the caller already wrote level 1. */
int dccm_vdiff_backward_device(dccm_vdiff *h,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    const double *level1, void *stream);

/* Columns that the reference-order forward solve handed to its plain-IEEE redo kernel since the handle was created
 * (columns whose operands left the range of the branch-free division: zero / denormal / non-finite values).  Results
 * are the IEEE operators' either way; the counter is for tests and diagnostics.  Synchronises the device. */
int dccm_vdiff_redo_total(dccm_vdiff *h, int64_t *columns);
/* the same two calls for columns [c0, c1) only (0-based, half open): the pieces of a latitude-slab pipeline */
int dccm_vdiff_forward_cols_device(dccm_vdiff *h,
    const double *xyr_MomFluxX, const double *xyr_MomFluxY, const double *xyr_HeatFlux,
    const double *xyrf_QMixFlux,
    const double *xyr_Press, const double *xyz_Exner, const double *xyr_Exner,
    const double *xyr_VirTemp, const double *xyz_Height,
    const double *xyr_VelDiffCoef, const double *xyr_TempDiffCoef, const double *xyr_QMixDiffCoef,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    double *xya_ImplCplCoef1, double *xya_ImplCplCoef2, int64_t c0, int64_t c1, void *stream);
int dccm_vdiff_backward_cols_device(dccm_vdiff *h,
    double *xyz_DUDt, double *xyz_DVDt, double *xyz_DTempDt, double *xyzf_DQMixDt,
    const double *level1, int64_t c0, int64_t c1, void *stream);

/* ------------------------------------------------------------------ ocean / sea-ice side
 * Element-wise work of ocn/dccm_ocn_mod.f90 either side of the remaps, on the device.
 * put: ice-surface selection (ref ocn/dccm_ocn_mod.f90:825-836) written straight into the O->S send layers
 *      o2s_bil = (SfcTemp ocean, SfcTemp ice), o2s_cons = (SIceCon, SfcAlbedo ocean, SfcAlbedo ice), row length ld.
 * get: from the 12 remapped S->O / S->I layers (row length ld; order as in dccm_sfc_exchange_device)
 *      FreshWtFlxS0 = ((Rain+Snow) - Evap)/DensFreshWater, FreshWtFlx0, sea-ice wind stress copies,
 *      SfcHFlxAO0 = ns + sr, DSfcHFlxAODTs (ref :978-993).
 * IceMaskMin, degC2K (DSIce) and DensFreshWater (DOGCM) belong to the external models. */
int dccm_ocn_put_assemble_device(int64_t n, const double *SeaSfcTemp, const double *SfcAlbedoAO,
                                 const double *SIceCon, const double *SIceSfcTempC, const double *SfcAlbedoAI,
                                 double IceMaskMin, double degC2K, double *o2s_bil, double *o2s_cons,
                                 int64_t ld, void *stream);
int dccm_ocn_get_assemble_device(int64_t n, const double *o_recv, int64_t ld, double DensFreshWater,
                                 double *FreshWtFlxS0, double *FreshWtFlx0, double *WindStressXAI,
                                 double *WindStressYAI, double *SfcHFlxAO0, double *DSfcHFlxAODTs, void *stream);

/* Jcup RECV_MODE='AVG' (ref ocn/dccm_ocn_mod.f90:652-672): time mean over the coupling interval of what the
 * surface component put, kept on the device.  accumulate: acc = x (first != 0) or acc = acc + x;
 * finish: acc = acc / count.  Jcup itself is not vendored in the reference tree; this is its documented
 * behaviour restated (parity unpinned), the arithmetic is pinned by the oracle (orc_avg_*). */
int dccm_avg_accumulate_device(double *acc, const double *x, int64_t n, int first, void *stream);
int dccm_avg_finish_device(double *acc, int64_t n, int count, void *stream);

/* ------------------------------------------------------------------ atmosphere side
 * dcpam_StoreAtmSurfFlxInfo, ref atm/dcpam_main_mod.f90:1040-1114: surface fluxes corrected with the level-1
 * tendencies of the implicit solve, element-wise over n columns (all pointers device, length n, none NULL).
 * DQVapSatDTempOnLiq / OnSol come from DCPAM's `saturate` module (external); the snow-fraction blend is done here. */
typedef struct dccm_atm_sfcflx {
    /* in */
    const double *SurfMomFluxX, *SurfMomFluxY, *SurfVelTransCoef, *SurfTempTransCoef, *SurfQVapTransCoef, *SurfHumidCoef;
    const double *DUDt1, *DVDt1, *DTempDtVDiff1, *DQVapDt1;     /* level 1 of the tendencies (IndexH2OVap for q) */
    const double *HeatFlux0, *QVapFlux0;                        /* level 0 of xyr_HeatFlux, xyrf_QMixFlux(:,:,0,IndexH2OVap) */
    const double *ExnerR0, *ExnerZ1, *TempN1, *DSurfTempDt;
    const double *SnowFrac, *DQVapSatDTempOnLiq, *DQVapSatDTempOnSol;
    const double *RadLDwFlux0, *RadLUwFlux0, *RadSDwFlux0, *RadSUwFlux0;
    const double *DelRadLDwFlux00, *DelRadLDwFlux01, *DelRadLUwFlux00, *DelRadLUwFlux01;
    /* out */
    double *TauXAtm, *TauYAtm, *SensAtm, *LatentAtm, *LDWRFlxAtm, *LUWRFlxAtm, *SDWRFlxAtm, *SUWRFlxAtm;
    double *SurfAirTemp, *DSurfLatentFlxDTs, *DSurfHFlxDTs;
} dccm_atm_sfcflx;
int dccm_atm_store_surf_flx_device(int64_t n, const dccm_atm_sfcflx *f, double LatentHeat, double CpDry,
                                   double DelTime, void *stream);

/* Atmosphere get side, ref atm/dccm_atm_mod.f90:823-836: from the 9 remapped S->A layers (row length ld, order as in
 * dccm_sfc_exchange_device's s2a) the surface temperature handed to the AGCM, xy_SfcTemp = (xy_LUwRFlx/StB)**0.25
 * (:831), and -- where the pointers are not NULL -- copies of SfcAlbedo (:827), SenHFlx -> xy_SurfHeatFlux (:825) and
 * QVapMFlx -> xyf_SurfQMixFlux(:,:,IndexH2OVap) (:826, :836).  Level 1 of the tendencies (:832-835) is the `level1`
 * argument of dccm_vdiff_backward_device.  StB is the glue's own constant (ref :789). */
int dccm_atm_get_assemble_device(int64_t n, const double *a_recv, int64_t ld, double StB, double *SfcTemp,
                                 double *SfcAlbedo, double *SurfHeatFlux, double *SurfH2OVapFlux, void *stream);

/* Legacy 2-component mode, atmosphere get side (ref atm/mod_atm.f90:740-775 + dcpam_UpdateSurfaceProperties,
 * atm/dcpam_main_mod.f90:1003-1031): o2a_recv = the 4 remapped O->A layers (SfcTemp**4, SfcAlbedo, SfcEngyFlxMod |
 * SfcSnow; row length ld).  SurfTemp = recv**0.25, SurfAlbedo = recv, SurfSnow = 1e3*recv, and level 1 of the
 * temperature (TempB1, in place) += SfcEngyFlxMod*cycle_sec / (Press0 - Press1) * Grav / CpDry, Press0 / Press1 being
 * xyr_Press at half levels 0 and 1 (DCPAM's AuxVars, external). */
int dccm_atm_legacy_get_assemble_device(int64_t n, const double *o2a_recv, int64_t ld, double cycle_sec,
                                        double Grav, double CpDry, const double *Press0, const double *Press1,
                                        double *SurfTemp, double *SurfAlbedo, double *SurfSnow, double *TempB1,
                                        void *stream);

#ifdef __cplusplus
}
#endif
#endif
