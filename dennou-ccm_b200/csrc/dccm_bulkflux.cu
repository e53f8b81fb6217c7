// dccm_bulkflux.cu -- K2: stand-alone bulk-flux kernel, one column per thread.
// Replaces DSFCM_Util_SfcBulkFlux_Get (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:108-439).
#include <cuda_runtime.h>

#include <algorithm>

#include "dccm_bulkflux.cuh"
#include "dccm_common.h"

using namespace dccm;

namespace {

struct BulkArgs {
    dccm_sfc_fields f;
    int nx, ny, ld;
    int64_t off, ss;
    double sig1;
};

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) bulkflux_kernel(const BulkArgs a)
{
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (int64_t)a.nx * a.ny) return;
    const int j = (int)(t / a.nx);
    const int i = (int)(t - (int64_t)j * a.nx);
    const int64_t c = a.off + i + (int64_t)a.ld * j;
    const int64_t ss = a.ss;
    const dccm_sfc_fields &f = a.f;

    BulkIn in;
    in.WindU = f.WindU[c]; in.WindV = f.WindV[c]; in.SfcAirTemp = f.SfcAirTemp[c]; in.QVap1 = f.QVap1[c];
    in.SDwRFlx = f.SDwRFlx[c]; in.LDwRFlx = f.LDwRFlx[c];
#pragma unroll
    for (int k = 0; k < 4; k++) { in.Coef1[k] = f.ImplCplCoef1[c + k * ss]; in.Coef2[k] = f.ImplCplCoef2[c + k * ss]; }
#pragma unroll
    for (int n = 0; n < 2; n++) { in.SfcTemp[n] = f.SfcTemp[c + n * ss]; in.SfcAlbedo[n] = f.SfcAlbedo[c + n * ss]; }
    in.SIceCon = f.SIceCon[c];
    in.SfcHeight = f.SfcHeight ? f.SfcHeight[c] : 0.0;
    in.SfcPress = f.SfcPress[c];

    BulkOut o;
    bulk_column(in, a.sig1, o);

#define ST3(ptr, v)                                                            \
    if (f.ptr) { f.ptr[c] = o.v[0]; f.ptr[c + ss] = o.v[1]; f.ptr[c + 2 * ss] = o.v[2]; }
#define ST2(ptr, v)                                                            \
    if (f.ptr) { f.ptr[c] = o.v[0]; f.ptr[c + ss] = o.v[1]; }
    ST3(WindStressX, WindStressX) ST3(WindStressY, WindStressY) ST3(SenHFlx, SenHFlx)
    ST3(QVapMFlx, QVapMFlx) ST3(LatHFlx, LatHFlx)
    ST3(SfcVelTransCoef, VelTC) ST3(SfcTempTransCoef, TempTC) ST3(SfcQVapTransCoef, QVapTC)
    ST3(SUwRFlx, SUwRFlx) ST3(LUwRFlx, LUwRFlx)
    ST2(SfcHFlx_ns, HFlx_ns) ST2(SfcHFlx_sr, HFlx_sr) ST2(DSfcHFlxDTs, DHFlxDTs)
#undef ST3
#undef ST2
    if (f.DelVarImplCPL) {
#pragma unroll
        for (int k = 0; k < 4; k++) f.DelVarImplCPL[c + k * ss] = o.Del[k];
    }
    f.SfcTemp[c + 2 * ss] = o.SfcTemp3;
    f.SfcAlbedo[c + 2 * ss] = o.SfcAlbedo3;
}

thread_local DevBuf g_buf;   // scratch of the host entry point (per calling thread)

// FastArith against the plain operators, element by element: counts[0] = accepted quotients / reciprocals /
// roots whose bits differ from `a/b`, `1.0/b`, `sqrt(|a|)` (must be 0), counts[1] = operations whose fast path
// was not acceptable (they take the IEEE re-evaluation in the product kernels).
__global__ void fast_arith_selftest_kernel(const double *a, const double *b, int64_t n, unsigned long long *counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = a[i], y = b[i];
    unsigned long long bad = 0, rej = 0;
    { FastArith f; const double q = f.div(x, y);
      if (!f.good()) rej++; else if (__double_as_longlong(q) != __double_as_longlong(x / y)) bad++; }
    { FastArith f; const double q = f.div_by(x, y, FastArith::prep(y));       // shared-reciprocal form of the division
      if (!f.good()) rej++; else if (__double_as_longlong(q) != __double_as_longlong(x / y)) bad++; }
    { FastArith f; const double q = f.rcp(y);
      if (!f.good()) rej++; else if (__double_as_longlong(q) != __double_as_longlong(1.0 / y)) bad++; }
    { FastArith f; const double q = f.root(fabs(x));
      if (!f.good()) rej++; else if (__double_as_longlong(q) != __double_as_longlong(sqrt(fabs(x)))) bad++; }
    if (bad) atomicAdd(&counts[0], bad);
    if (rej) atomicAdd(&counts[1], rej);
}

// dccm_pmath.cuh element by element, through the product's own arithmetic (FastArith, IEEE redo when rejected)
__global__ void pmath_selftest_kernel(int which, const double *x, int64_t n, double y, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r;
    if (which == 0) r = pexp(v);
    else if (which == 3) r = pfourth_root(v);
    else {
        FastArith f;
        r = (which == 1) ? plog(v, f) : ppow(v, y, f);
        if (!f.good()) { IeeeArith g; r = (which == 1) ? plog(v, g) : ppow(v, y, g); }
    }
    out[i] = r;
}

}  // namespace

extern "C" int dccm_selftest_pmath_device(int which, const double *d_x, int64_t n, double y, double *d_out)
{
    if (which < 0 || which > 3 || n < 0) return fail(DCCM_ERR_ARG, "dccm_selftest_pmath: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    if (n == 0) return DCCM_OK;
    pmath_selftest_kernel<<<(unsigned)((n + 255) / 256), 256>>>(which, d_x, n, y, d_out);
    DCCM_CUDA_TRY(cudaGetLastError());
    DCCM_CUDA_TRY(cudaDeviceSynchronize());
    return DCCM_OK;
}

extern "C" int dccm_bulkflux_device(int nx, int ny, int ld, int64_t off, int64_t slot_stride,
                                    const dccm_sfc_fields *f, double sig1, void *stream)
{
    NvtxRange nvtx("dccm_bulkflux_device");
    if (!f) return fail(DCCM_ERR_ARG, "dccm_bulkflux: null field table");
    if (nx < 1 || ny < 1 || ld < nx) return fail(DCCM_ERR_ARG, "dccm_bulkflux: bad extents nx=%d ny=%d ld=%d", nx, ny, ld);
    if (!f->WindU || !f->WindV || !f->SfcAirTemp || !f->QVap1 || !f->SDwRFlx || !f->LDwRFlx || !f->ImplCplCoef1 ||
        !f->ImplCplCoef2 || !f->SfcTemp || !f->SfcAlbedo || !f->SIceCon || !f->SfcPress)
        return fail(DCCM_ERR_ARG, "dccm_bulkflux: a required input pointer is NULL");
    int rc = ensure_device();
    if (rc) return rc;
    BulkArgs a;
    a.f = *f; a.nx = nx; a.ny = ny; a.ld = ld; a.off = off; a.ss = slot_stride; a.sig1 = sig1;
    const int64_t n = (int64_t)nx * ny;
    const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
    bulkflux_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

extern "C" int dccm_selftest_fast_arith_device(const double *d_a, const double *d_b, int64_t n,
                                               int64_t *mismatches, int64_t *rejected)
{
    int rc = ensure_device();
    if (rc) return rc;
    unsigned long long *d_counts = nullptr, h[2] = {0, 0};
    DCCM_CUDA_TRY(cudaMalloc(&d_counts, sizeof h));
    DCCM_CUDA_TRY(cudaMemset(d_counts, 0, sizeof h));
    fast_arith_selftest_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_a, d_b, n, d_counts);
    cudaError_t e = cudaMemcpy(h, d_counts, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d_counts);
    DCCM_CUDA_TRY(e);
    *mismatches = (int64_t)h[0]; *rejected = (int64_t)h[1];
    return DCCM_OK;
}

// Host (drop-in) form: explicit-shape (IA,JA[,n]) arrays with halo 1, interior IS:IE,JS:JE
// (ref sfc/DSFCM_Admin_Grid_mod.f90:39-50).  Halo cells of the caller's arrays are untouched:
// only the interior rows are moved back (2-D copies), as the reference only writes IS:IE,JS:JE.
extern "C" int dccm_bulkflux_get_host(int IA, int JA,
    double *WSX, double *WSY, double *SenH, double *QVapM, double *LatH,
    double *VelTC, double *TempTC, double *QVapTC, double *Del,
    double *SUw, double *LUw, double *HFns, double *HFsr, double *DHFDTs,
    const double *WindU, const double *WindV, const double *SfcAirTemp, const double *QVap1,
    const double *SDw, const double *LDw, const double *Coef1, const double *Coef2,
    double *SfcTemp, double *SfcAlbedo, const double *SIceCon,
    const double *Sig1Info, const double *SfcHeight, const double *SfcPress)
{
    NvtxRange nvtx("dccm_bulkflux_get_host");
    if (IA < 3 || JA < 3) return fail(DCCM_ERR_ARG, "dccm_bulkflux_get: IA, JA must include the halo (>= 3)");
    int rc = ensure_device();
    if (rc) return rc;
    const size_t N2 = (size_t)IA * JA;
    // device mirror: 14 outputs (13 x 3 slots + Del x 4) + inputs (9 x 1 + 2 x 4 + 2 x 3)
    const size_t n_out = 13 * 3 + 4, n_in = 9 + 8 + 6;
    rc = g_buf.reserve(sizeof(double) * N2 * (n_out + n_in));
    if (rc) return rc;
    double *d = g_buf.as<double>();
    size_t pos = 0;
    auto take = [&](size_t slots) { double *p = d + pos * N2; pos += slots; return p; };
    dccm_sfc_fields f;
    f.WindStressX = take(3); f.WindStressY = take(3); f.SenHFlx = take(3); f.QVapMFlx = take(3); f.LatHFlx = take(3);
    f.SfcVelTransCoef = take(3); f.SfcTempTransCoef = take(3); f.SfcQVapTransCoef = take(3);
    f.DelVarImplCPL = take(4);
    f.SUwRFlx = take(3); f.LUwRFlx = take(3); f.SfcHFlx_ns = take(3); f.SfcHFlx_sr = take(3); f.DSfcHFlxDTs = take(3);
    double *dWindU = take(1), *dWindV = take(1), *dT1 = take(1), *dQ1 = take(1), *dSDw = take(1), *dLDw = take(1);
    double *dC1 = take(4), *dC2 = take(4), *dTs = take(3), *dAl = take(3), *dIce = take(1), *dH = take(1), *dPs = take(1);
    f.WindU = dWindU; f.WindV = dWindV; f.SfcAirTemp = dT1; f.QVap1 = dQ1; f.SDwRFlx = dSDw; f.LDwRFlx = dLDw;
    f.ImplCplCoef1 = dC1; f.ImplCplCoef2 = dC2; f.SfcTemp = dTs; f.SfcAlbedo = dAl; f.SIceCon = dIce;
    f.SfcHeight = dH; f.SfcPress = dPs;

    // Interior rows move in chunks: chunk j+1 on its way in (H2D stream) while the kernel runs on chunk j and
    // chunk j-1 is on its way out (D2H stream).  Inputs are whole rows of the (IA,JA) slots (contiguous), outputs
    // the interior columns only -- halo cells of the caller's arrays are never written, as in the reference.
    static thread_local cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    for (auto &ps : pipe)
        if (!ps) DCCM_CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    cudaStream_t sin = pipe[0], sk = pipe[1], sout = pipe[2];
    const int nx = IA - 2, ny = JA - 2;
    const char *env = getenv("DCCM_HOST_CHUNKS");
    int nchunk = env ? std::max(1, atoi(env)) : ((size_t)nx * ny >= (1u << 19) ? 12 : 1);
    nchunk = std::min(nchunk, ny);
    std::vector<cudaEvent_t> ev(2 * (size_t)nchunk, nullptr);
    struct EvGuard { std::vector<cudaEvent_t> &e; ~EvGuard() { for (auto x : e) if (x) cudaEventDestroy(x); } } guard{ev};
    for (auto &e : ev) DCCM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    struct { double *d; const double *h; int slots; } ins[] = {
        {dWindU, WindU, 1}, {dWindV, WindV, 1}, {dT1, SfcAirTemp, 1}, {dQ1, QVap1, 1}, {dSDw, SDw, 1}, {dLDw, LDw, 1},
        {dC1, Coef1, 4}, {dC2, Coef2, 4}, {dTs, SfcTemp, 2}, {dAl, SfcAlbedo, 2}, {dIce, SIceCon, 1}, {dH, SfcHeight, 1},
        {dPs, SfcPress, 1}};
    struct { double *h; const double *d; int first, last; } outs[] = {
        {WSX, f.WindStressX, 0, 2}, {WSY, f.WindStressY, 0, 2}, {SenH, f.SenHFlx, 0, 2}, {QVapM, f.QVapMFlx, 0, 2},
        {LatH, f.LatHFlx, 0, 2}, {VelTC, f.SfcVelTransCoef, 0, 2}, {TempTC, f.SfcTempTransCoef, 0, 2},
        {QVapTC, f.SfcQVapTransCoef, 0, 2}, {Del, f.DelVarImplCPL, 0, 3}, {SUw, f.SUwRFlx, 0, 2}, {LUw, f.LUwRFlx, 0, 2},
        {HFns, f.SfcHFlx_ns, 0, 1}, {HFsr, f.SfcHFlx_sr, 0, 1}, {DHFDTs, f.DSfcHFlxDTs, 0, 1},
        {SfcTemp, f.SfcTemp, 2, 2}, {SfcAlbedo, f.SfcAlbedo, 2, 2}};
    const int per = (ny + nchunk - 1) / nchunk;
    for (int c = 0; c < nchunk; c++) {
        const int j0 = c * per, j1 = std::min(ny, j0 + per);        // interior rows [j0, j1) = array rows j0+1 .. j1
        if (j0 >= j1) continue;
        const size_t row0 = (size_t)(j0 + 1) * IA, nrow = (size_t)(j1 - j0) * IA;
        for (auto &i : ins)
            for (int sl = 0; sl < i.slots; sl++)
                DCCM_CUDA_TRY(cudaMemcpyAsync(i.d + sl * N2 + row0, i.h + sl * N2 + row0, sizeof(double) * nrow,
                                              cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(cudaEventRecord(ev[2 * c], sin));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sk, ev[2 * c], 0));
        rc = dccm_bulkflux_device(nx, j1 - j0, IA, (int64_t)row0 + 1, (int64_t)N2, &f, Sig1Info[0], sk);
        if (rc) return rc;
        DCCM_CUDA_TRY(cudaEventRecord(ev[2 * c + 1], sk));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sout, ev[2 * c + 1], 0));
        for (auto &o : outs)
            for (int sl = o.first; sl <= o.last; sl++) {
                const size_t at = (size_t)sl * N2 + row0 + 1;
                DCCM_CUDA_TRY(cudaMemcpy2DAsync(o.h + at, sizeof(double) * IA, o.d + at, sizeof(double) * IA,
                                                sizeof(double) * nx, j1 - j0, cudaMemcpyDeviceToHost, sout));
            }
    }
    DCCM_CUDA_TRY(cudaStreamSynchronize(sout));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sk));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sin));
    return DCCM_OK;
}
