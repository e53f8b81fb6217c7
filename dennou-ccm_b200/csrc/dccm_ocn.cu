// dccm_ocn.cu -- the ocean / sea-ice component's element-wise work either side of the remaps
// (SURVEY.md 8f rank 3), kept on the device so the O->S send layers and the S->O results never
// visit the host:
//   put side: ice-surface selection by IceMaskMin           ref ocn/dccm_ocn_mod.f90:825-836, :847-851
//   get side: fresh-water flux, net ocean heat flux, copies  ref ocn/dccm_ocn_mod.f90:978-993
// DensFreshWater, IceMaskMin and degC2K come from DOGCM / DSIce (external models): parameters here.
#include <cuda_runtime.h>

#include "dccm_common.h"

using namespace dccm;

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
ocn_put_kernel(int64_t n, const double *__restrict__ SeaSfcTemp, const double *__restrict__ AlbAO,
               const double *__restrict__ SIceCon, const double *__restrict__ SIceSfcTempC,
               const double *__restrict__ AlbAI, double IceMaskMin, double degC2K,
               double *__restrict__ bil, double *__restrict__ cons, int64_t ld)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const double ice = SIceCon[c], ts = SeaSfcTemp[c], ao = AlbAO[c];
    double ti, ai;
    if (ice >= IceMaskMin) { ti = SIceSfcTempC[c] + degC2K; ai = AlbAI[c]; }      // :828-830
    else                   { ti = ts;                       ai = ao; }            // :831-833
    bil[c] = ts; bil[c + ld] = ti;                           // o2s_SfcTemp, i2s_SfcTemp      (:847,:849)
    cons[c] = ice; cons[c + ld] = ao; cons[c + 2 * ld] = ai; // i2s_SIceCon, o2s/i2s_SfcAlbedo (:848,:850-851)
}

__global__ void __launch_bounds__(kThreads)
ocn_get_kernel(int64_t n, const double *__restrict__ r, int64_t ld, double DensFreshWater,
               double *__restrict__ FreshWtFlxS0, double *__restrict__ FreshWtFlx0,
               double *__restrict__ WSXAI, double *__restrict__ WSYAI,
               double *__restrict__ SfcHFlxAO0, double *__restrict__ DSfcHFlxAODTs)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    // S->O layers (exchange.py S2O_CONS / S2O_BIL): 0 ns, 1 sr, 2 snow, 3 rain, 4 evap, 5 taux, 6 tauy, 10 dF/dTs
    const double ns = r[c], sr = r[c + ld], snow = r[c + 2 * ld], rain = r[c + 3 * ld], evap = r[c + 4 * ld];
    const double fw = ((rain + snow) - evap) / DensFreshWater;                    // :980-983
    FreshWtFlxS0[c] = fw;
    FreshWtFlx0[c] = fw;                                                          // :984
    WSXAI[c] = r[c + 5 * ld];                                                     // :987-988
    WSYAI[c] = r[c + 6 * ld];
    SfcHFlxAO0[c] = ns + sr;                                                      // :990
    DSfcHFlxAODTs[c] = r[c + 10 * ld];                                            // :991
}
// Jcup RECV_MODE='AVG' (ref ocn/dccm_ocn_mod.f90:652-672: every S->O / S->I variable): the receiver sees the
// mean of what the sender put during the coupling interval.  acc = x on the first put of an interval,
// acc = acc + x afterwards, acc = acc / count when the interval closes; layers are rows of length ld.
__global__ void __launch_bounds__(kThreads)
avg_accumulate_kernel(double *__restrict__ acc, const double *__restrict__ x, int64_t n, int first)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    acc[c] = first ? x[c] : acc[c] + x[c];
}

__global__ void __launch_bounds__(kThreads) avg_finish_kernel(double *__restrict__ acc, int64_t n, double count)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    acc[c] = acc[c] / count;
}
}  // namespace

extern "C" int dccm_avg_accumulate_device(double *acc, const double *x, int64_t n, int first, void *stream)
{
    if (!acc || !x || n < 1) return fail(DCCM_ERR_ARG, "dccm_avg_accumulate: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    avg_accumulate_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        acc, x, n, first);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_avg_finish_device(double *acc, int64_t n, int count, void *stream)
{
    if (!acc || n < 1 || count < 1) return fail(DCCM_ERR_ARG, "dccm_avg_finish: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    avg_finish_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        acc, n, (double)count);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_ocn_put_assemble_device(int64_t n, const double *SeaSfcTemp, const double *SfcAlbedoAO,
                                            const double *SIceCon, const double *SIceSfcTempC, const double *SfcAlbedoAI,
                                            double IceMaskMin, double degC2K, double *o2s_bil, double *o2s_cons,
                                            int64_t ld, void *stream)
{
    NvtxRange nvtx("dccm_ocn_put_assemble_device");
    if (n < 1 || ld < n) return fail(DCCM_ERR_ARG, "dccm_ocn_put_assemble: bad extents");
    if (!SeaSfcTemp || !SfcAlbedoAO || !SIceCon || !SIceSfcTempC || !SfcAlbedoAI || !o2s_bil || !o2s_cons)
        return fail(DCCM_ERR_ARG, "dccm_ocn_put_assemble: null pointer");
    int rc = ensure_device();
    if (rc) return rc;
    ocn_put_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        n, SeaSfcTemp, SfcAlbedoAO, SIceCon, SIceSfcTempC, SfcAlbedoAI, IceMaskMin, degC2K, o2s_bil, o2s_cons, ld);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_ocn_get_assemble_device(int64_t n, const double *o_recv, int64_t ld, double DensFreshWater,
                                            double *FreshWtFlxS0, double *FreshWtFlx0, double *WindStressXAI,
                                            double *WindStressYAI, double *SfcHFlxAO0, double *DSfcHFlxAODTs, void *stream)
{
    NvtxRange nvtx("dccm_ocn_get_assemble_device");
    if (n < 1 || ld < n) return fail(DCCM_ERR_ARG, "dccm_ocn_get_assemble: bad extents");
    if (!o_recv || !FreshWtFlxS0 || !FreshWtFlx0 || !WindStressXAI || !WindStressYAI || !SfcHFlxAO0 || !DSfcHFlxAODTs)
        return fail(DCCM_ERR_ARG, "dccm_ocn_get_assemble: null pointer");
    if (!(DensFreshWater > 0.0)) return fail(DCCM_ERR_ARG, "dccm_ocn_get_assemble: DensFreshWater must be positive");
    int rc = ensure_device();
    if (rc) return rc;
    ocn_get_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        n, o_recv, ld, DensFreshWater, FreshWtFlxS0, FreshWtFlx0, WindStressXAI, WindStressYAI, SfcHFlxAO0, DSfcHFlxAODTs);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}
