// dccm_atm.cu -- the atmosphere component's element-wise surface-flux bookkeeping after the backward
// solve: dcpam_StoreAtmSurfFlxInfo, ref atm/dcpam_main_mod.f90:1040-1114.
// Level-1 tendencies of the implicit solve correct the explicit surface fluxes; the results are what the
// legacy 2-component glue ships to the ocean (ref atm/mod_atm.f90:653-667).  One thread per column, every
// operand read once; operation order as written in the reference (the library is built with -fmad=false).
// The saturation functions (xy_CalcDQVapSatDTempOnLiq/OnSol, DCPAM `saturate`) belong to the external model:
// their values are inputs, the snow-fraction blend of :1068-1070 is done here.
#include <cuda_runtime.h>

#include "dccm_common.h"
#include "dccm_pmath.cuh"

using namespace dccm;

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
atm_store_surf_flx_kernel(int64_t n, const dccm_atm_sfcflx f, double LatentHeat, double CpDry, double delta_t)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const double dTs = f.DSurfTempDt[c], dT1 = f.DTempDtVDiff1[c];
    const double ex0 = f.ExnerR0[c], ex1 = f.ExnerZ1[c];
    const double velTC = f.SurfVelTransCoef[c], tempTC = f.SurfTempTransCoef[c], qTC = f.SurfQVapTransCoef[c];
    const double humid = f.SurfHumidCoef[c];
    const double snow = f.SnowFrac[c];
    const double dqsat = (1.0 - snow) * f.DQVapSatDTempOnLiq[c] + snow * f.DQVapSatDTempOnSol[c];   // :1068-1070

    f.TauXAtm[c] = f.SurfMomFluxX[c] - velTC * f.DUDt1[c] * 2.0 * delta_t;                           // :1078
    f.TauYAtm[c] = f.SurfMomFluxY[c] - velTC * f.DVDt1[c] * 2.0 * delta_t;                           // :1079
    f.SensAtm[c] = f.HeatFlux0[c] - CpDry * ex0 * tempTC * (dT1 / ex1 - dTs / ex0) * 2.0 * delta_t;  // :1081-1084
    f.LatentAtm[c] = LatentHeat * (f.QVapFlux0[c]                                                    // :1085-1090
                                   - humid * qTC * (f.DQVapDt1[c] - dqsat * dTs) * 2.0 * delta_t);
    f.LDWRFlxAtm[c] = f.RadLDwFlux0[c] + 2.0 * delta_t * (dTs * f.DelRadLDwFlux00[c] + dT1 * f.DelRadLDwFlux01[c]);   // :1092-1095
    f.LUWRFlxAtm[c] = f.RadLUwFlux0[c] + 2.0 * delta_t * (dTs * f.DelRadLUwFlux00[c] + dT1 * f.DelRadLUwFlux01[c]);   // :1096-1099
    f.SDWRFlxAtm[c] = f.RadSDwFlux0[c];                                                              // :1101
    f.SUWRFlxAtm[c] = f.RadSUwFlux0[c];                                                              // :1102
    f.SurfAirTemp[c] = ex0 / ex1 * f.TempN1[c];                                                      // :1106
    const double dlat = LatentHeat * humid * qTC * dqsat;                                            // :1107
    f.DSurfLatentFlxDTs[c] = dlat;
    f.DSurfHFlxDTs[c] = CpDry * tempTC + dlat - f.DelRadLDwFlux00[c];                                // :1108-1112
}

// Atmosphere get side after the S->A remaps (ref atm/dccm_atm_mod.f90:823-836): the surface temperature the AGCM is
// handed is the radiative one, (LUwRFlx / StB)**0.25 of the composite upward long-wave flux; the other gets are
// the remapped layers themselves (albedo, fluxes) or go to level 1 of the tendencies (the backward solve's level1).
__global__ void __launch_bounds__(kThreads)
atm_get_kernel(int64_t n, const double *__restrict__ a_recv, int64_t ld, double StB, double *__restrict__ SfcTemp,
               double *__restrict__ SfcAlbedo, double *__restrict__ SurfHeatFlux, double *__restrict__ SurfH2OVapFlux)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    // S->A layers (exchange.py S2A_CONS / S2A_BIL): 0 LUwRFlx, 1 SUwRFlx, 2 SenHFlx, 3 QVapMFlx, 4 SfcAlbedo, 5..8 DelVarImplCPL
    SfcTemp[c] = pfourth_root(a_recv[c] / StB);                                                         // :831
    if (SfcAlbedo) SfcAlbedo[c] = a_recv[c + 4 * ld];                                                // :827
    if (SurfHeatFlux) SurfHeatFlux[c] = a_recv[c + 2 * ld];                                          // :825
    if (SurfH2OVapFlux) SurfH2OVapFlux[c] = a_recv[c + 3 * ld];                                      // :826, :836
}

// Legacy 2-component mode, atmosphere get side (ref atm/mod_atm.f90:740-775, atm/dcpam_main_mod.f90:1003-1031): the four
// remapped O->A layers (SfcTemp**4, SfcAlbedo, SfcEngyFlxMod | SfcSnow) become the AGCM's surface temperature (fourth
// root), albedo, snow (x 1e3) and a correction of the lowest-level temperature by the ocean's energy-flux residual
// accumulated over the coupling cycle.
__global__ void __launch_bounds__(kThreads)
atm_legacy_get_kernel(int64_t n, const double *__restrict__ r, int64_t ld, double cycle_sec, double Grav, double CpDry,
                      const double *__restrict__ Press0, const double *__restrict__ Press1,
                      double *__restrict__ SurfTemp, double *__restrict__ SurfAlbedo, double *__restrict__ SurfSnow,
                      double *__restrict__ TempB1)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    SurfTemp[c] = pfourth_root(r[c]);                                                                   // mod_atm.f90:743
    SurfAlbedo[c] = r[c + ld];                                                                       // :745-746, dcpam_main_mod.f90:1017
    SurfSnow[c] = 1e3 * r[c + 3 * ld];                                                               // mod_atm.f90:772
    const double mod = r[c + 2 * ld] * cycle_sec;                                                    // :773
    TempB1[c] = TempB1[c] + (mod - 0.0) / (Press0[c] - Press1[c]) * Grav / CpDry;                    // dcpam_main_mod.f90:1026-1028
}
}  // namespace

extern "C" int dccm_atm_legacy_get_assemble_device(int64_t n, const double *o2a_recv, int64_t ld, double cycle_sec,
                                                   double Grav, double CpDry, const double *Press0, const double *Press1,
                                                   double *SurfTemp, double *SurfAlbedo, double *SurfSnow, double *TempB1,
                                                   void *stream)
{
    if (!o2a_recv || !Press0 || !Press1 || !SurfTemp || !SurfAlbedo || !SurfSnow || !TempB1)
        return fail(DCCM_ERR_ARG, "dccm_atm_legacy_get_assemble: null buffer");
    if (n < 1 || ld < n) return fail(DCCM_ERR_ARG, "dccm_atm_legacy_get_assemble: need 1 <= n <= ld");
    int rc = ensure_device();
    if (rc) return rc;
    atm_legacy_get_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        n, o2a_recv, ld, cycle_sec, Grav, CpDry, Press0, Press1, SurfTemp, SurfAlbedo, SurfSnow, TempB1);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_atm_get_assemble_device(int64_t n, const double *a_recv, int64_t ld, double StB, double *SfcTemp,
                                            double *SfcAlbedo, double *SurfHeatFlux, double *SurfH2OVapFlux, void *stream)
{
    NvtxRange nvtx("dccm_atm_get_assemble_device");
    if (!a_recv || !SfcTemp) return fail(DCCM_ERR_ARG, "dccm_atm_get_assemble: null buffer");
    if (n < 1 || ld < n) return fail(DCCM_ERR_ARG, "dccm_atm_get_assemble: need 1 <= n <= ld");
    if (!(StB > 0.0)) return fail(DCCM_ERR_ARG, "dccm_atm_get_assemble: StB must be positive");
    int rc = ensure_device();
    if (rc) return rc;
    atm_get_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        n, a_recv, ld, StB, SfcTemp, SfcAlbedo, SurfHeatFlux, SurfH2OVapFlux);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_atm_store_surf_flx_device(int64_t n, const dccm_atm_sfcflx *f, double LatentHeat, double CpDry,
                                              double DelTime, void *stream)
{
    NvtxRange nvtx("dccm_atm_store_surf_flx_device");
    if (!f) return fail(DCCM_ERR_ARG, "dccm_atm_store_surf_flx: null field table");
    if (n < 1) return fail(DCCM_ERR_ARG, "dccm_atm_store_surf_flx: n must be >= 1");
    const void *const *p = reinterpret_cast<const void *const *>(f);
    for (size_t i = 0; i < sizeof(*f) / sizeof(void *); i++)
        if (!p[i]) return fail(DCCM_ERR_ARG, "dccm_atm_store_surf_flx: field pointer %zu is NULL", i);
    int rc = ensure_device();
    if (rc) return rc;
    atm_store_surf_flx_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0,
                                reinterpret_cast<cudaStream_t>(stream)>>>(n, *f, LatentHeat, CpDry, DelTime);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}
