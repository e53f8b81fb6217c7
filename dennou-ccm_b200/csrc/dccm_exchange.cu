// dccm_exchange.cu -- K12: the surface component's whole coupling step in ONE kernel.
//
// For every cell of the surface exchange grid one thread
//   1. gathers the 17 atmosphere layers (13 through the bilinear A->S table, 4 through the
//      conservative one) and the 5 ocean / sea-ice layers (2 bilinear, 3 conservative) straight
//      from the ATM- and OCN-grid send buffers -- the remap of
//      ref common/interpolation_data_latlon_mod.f90:293-302 with the field grouping of
//      ref sfc/dccm_sfc_mod.f90:449-466 -- accumulating each layer in table order with separate
//      IEEE multiply/add, i.e. bit-identical to K1 run layer group by layer group;
//   2. evaluates the bulk flux + implicit surface update in registers (dccm_bulkflux.cuh,
//      ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:194-415);
//   3. stores ONLY what the glue puts back to the coupler, already packed in send order
//      (ref sfc/dccm_sfc_mod.f90:764-784): 9 layers for the atmosphere, 12 for ocean / sea ice
//      (wind stress negated, rain / snow passed through).
// The 22 remapped input layers and the 21 unused bulk outputs never touch HBM, and the
// reference's unpack -> (IA,JA) halo arrays -> pack staging disappears.  Optionally the
// API-complete DSFCM arrays are stored as well (diagnostics / parity tests).
#include <cuda_runtime.h>

#include "dccm_bulkflux.cuh"
#include "dccm_remap_internal.h"

using namespace dccm;

namespace {

constexpr int kThreads = 128;

struct Csr {                       // one table: CSR (kind 0) or zonal stencil (kind 1)
    const int32_t *rowptr, *col;
    const double *w;
    int kind, nxs, nxd;
};

struct SfcArgs {
    Csr as_bil, as_cons, os_bil, os_cons;
    SrcSeg a2s_bil, a2s_cons, o2s_bil, o2s_cons;             // (13M, nA) (4M, nA) (2M, nO) (3M, nO)
    double *s2a, *s2o;                                        // (9M, nS) (12M, nS)
    dccm_sfc_fields full;                                     // optional API-complete outputs, slot stride M*nS
    int has_full;
    int64_t nA, nO, nS, sld;      // sld: cells per (layer, member) row of s2a / s2o (>= nS)
    int M;
    double sig1;
};

// One destination row of one table, D layers.  The row's (col, w) pairs are fetched CH at a
// time BEFORE any of the dependent source loads is issued, so a thread has up to CH*D gathers in
// flight instead of D (rows of these tables hold 1-4 entries); accumulation stays in table order.
template <int D, int CH>
__device__ __forceinline__ void gather(const Csr &t, int r, const SrcSeg &src, int64_t n_src,
                                       int M, int m, double (&acc)[D])
{
#pragma unroll
    for (int d = 0; d < D; d++) acc[d] = 0.0;
    const int64_t o0 = (int64_t)m * n_src;
    const int64_t lstride = (int64_t)M * n_src;
    if (t.kind == 1) {
        // zonal stencil: rowptr = per-latitude-row pointer, col = interleaved (di, jS) pairs
        const int jD = r / t.nxd, iD = r - jD * t.nxd;
        const int e0 = __ldg(&t.rowptr[jD]), e1 = __ldg(&t.rowptr[jD + 1]);
        for (int e = e0; e < e1; e++) {
            int i = iD + __ldg(&t.col[2 * e]);
            if (i >= t.nxs) i -= t.nxs;
            if (t.nxs == 1) i = 0;                   // axisymmetric source
            const double *p = src.at((int64_t)__ldg(&t.col[2 * e + 1]) * t.nxs + i) + o0;
            const double ww = __ldg(&t.w[e]);
#pragma unroll
            for (int d = 0; d < D; d++) acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(p + d * lstride), ww));
        }
        return;
    }
    const int k0 = __ldg(&t.rowptr[r]), k1 = __ldg(&t.rowptr[r + 1]);
    for (int kb = k0; kb < k1; kb += CH) {
        int c[CH];
        double w[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const bool on = kb + j < k1;
            c[j] = on ? __ldg(&t.col[kb + j]) : 0;
            w[j] = on ? __ldg(&t.w[kb + j]) : 0.0;
        }
        double v[CH][D];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            if (kb + j < k1) {
                const double *p = src.at(c[j]) + o0;
#pragma unroll
                for (int d = 0; d < D; d++) v[j][d] = __ldg(p + d * lstride);
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
            if (kb + j < k1) {
#pragma unroll
                for (int d = 0; d < D; d++) acc[d] = __dadd_rn(acc[d], __dmul_rn(v[j][d], w[j]));
            }
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) sfc_exchange_kernel(const SfcArgs a)
{
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nS = a.nS;
    if (t >= nS * a.M) return;
    const int m = (int)(t / nS);
    const int r = (int)(t - (int64_t)m * nS);
    const int M = a.M;

    double ab[13], ac[4], ob[2], oc[3];
    gather<13, 2>(a.as_bil, r, a.a2s_bil, a.nA, M, m, ab);
    gather<4, 4>(a.as_cons, r, a.a2s_cons, a.nA, M, m, ac);
    gather<2, 4>(a.os_bil, r, a.o2s_bil, a.nO, M, m, ob);
    gather<3, 4>(a.os_cons, r, a.o2s_cons, a.nO, M, m, oc);

    BulkIn in;
    in.WindU = ab[0]; in.WindV = ab[1]; in.SfcAirTemp = ab[2]; in.QVap1 = ab[3]; in.SfcPress = ab[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { in.Coef1[k] = ab[5 + k]; in.Coef2[k] = ab[9 + k]; }
    in.LDwRFlx = ac[0]; in.SDwRFlx = ac[1];
    const double rain = ac[2], snow = ac[3];
    in.SfcTemp[0] = ob[0]; in.SfcTemp[1] = ob[1];
    in.SIceCon = oc[0]; in.SfcAlbedo[0] = oc[1]; in.SfcAlbedo[1] = oc[2];
    in.SfcHeight = 0.0;                                   // ref sfc/dccm_sfc_mod.f90:885

    BulkOut o;
    bulk_column(in, a.sig1, o);

    // packed put-side layers, row = layer * M + member
    const int64_t ld = (int64_t)M * a.sld;                // one layer of all members
    double *pa = a.s2a + (int64_t)m * a.sld + r;
    pa[0 * ld] = o.LUwRFlx[2]; pa[1 * ld] = o.SUwRFlx[2]; pa[2 * ld] = o.SenHFlx[2]; pa[3 * ld] = o.QVapMFlx[2];
    pa[4 * ld] = o.SfcAlbedo3;
#pragma unroll
    for (int k = 0; k < 4; k++) pa[(5 + k) * ld] = o.Del[k];
    double *po = a.s2o + (int64_t)m * a.sld + r;
    po[0 * ld] = o.HFlx_ns[0]; po[1 * ld] = o.HFlx_sr[0]; po[2 * ld] = snow; po[3 * ld] = rain;
    po[4 * ld] = o.QVapMFlx[0]; po[5 * ld] = -o.WindStressX[2]; po[6 * ld] = -o.WindStressY[2];
    po[7 * ld] = o.HFlx_ns[1]; po[8 * ld] = o.HFlx_sr[1]; po[9 * ld] = o.QVapMFlx[1];
    po[10 * ld] = o.DHFlxDTs[0]; po[11 * ld] = o.DHFlxDTs[1];

    if (a.has_full) {
        const dccm_sfc_fields &f = a.full;
        const int64_t c = (int64_t)m * nS + r, ss = (int64_t)M * nS;
#define ST3(ptr, v) if (f.ptr) { f.ptr[c] = o.v[0]; f.ptr[c + ss] = o.v[1]; f.ptr[c + 2 * ss] = o.v[2]; }
#define ST2(ptr, v) if (f.ptr) { f.ptr[c] = o.v[0]; f.ptr[c + ss] = o.v[1]; }
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
        if (f.SfcTemp) { f.SfcTemp[c] = in.SfcTemp[0]; f.SfcTemp[c + ss] = in.SfcTemp[1]; f.SfcTemp[c + 2 * ss] = o.SfcTemp3; }
        if (f.SfcAlbedo) { f.SfcAlbedo[c] = in.SfcAlbedo[0]; f.SfcAlbedo[c + ss] = in.SfcAlbedo[1]; f.SfcAlbedo[c + 2 * ss] = o.SfcAlbedo3; }
    }
}

Csr csr_of(const dccm_remap *h)
{
    if (h->kind == 1) return Csr{h->d_zptr, h->d_zdj, h->d_zw, 1, h->nxs, h->nxd};
    return Csr{h->d_rowptr, h->d_col, h->d_w, 0, 0, 0};
}

}  // namespace

extern "C" int dccm_sfc_exchange_device(const dccm_remap *as_bil, const dccm_remap *as_cons,
                                        const dccm_remap *os_bil, const dccm_remap *os_cons,
                                        const double *a2s_bil, const double *a2s_cons,
                                        const double *o2s_bil, const double *o2s_cons,
                                        int members, double sig1, double *s2a, double *s2o, int64_t s_ld,
                                        const dccm_sfc_fields *full, void *stream)
{
    if (!a2s_bil || !a2s_cons || !o2s_bil || !o2s_cons) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: null buffer");
    dccm_src_seg a{a2s_bil, a2s_bil, a2s_bil, 0, INT64_MAX}, b{a2s_cons, a2s_cons, a2s_cons, 0, INT64_MAX};
    dccm_src_seg c{o2s_bil, o2s_bil, o2s_bil, 0, INT64_MAX}, d{o2s_cons, o2s_cons, o2s_cons, 0, INT64_MAX};
    return dccm_sfc_exchange_seg_device(as_bil, as_cons, os_bil, os_cons, &a, &b, &c, &d, 0, 0, members, sig1,
                                        s2a, s2o, s_ld, full, stream);
}

extern "C" int dccm_sfc_exchange_seg_device(const dccm_remap *as_bil, const dccm_remap *as_cons,
                                            const dccm_remap *os_bil, const dccm_remap *os_cons,
                                            const dccm_src_seg *sa2s_bil, const dccm_src_seg *sa2s_cons,
                                            const dccm_src_seg *so2s_bil, const dccm_src_seg *so2s_cons,
                                            int64_t a_ld, int64_t o_ld,
                                            int members, double sig1, double *s2a, double *s2o, int64_t s_ld,
                                            const dccm_sfc_fields *full, void *stream)
{
    if (!sa2s_bil || !sa2s_cons || !so2s_bil || !so2s_cons) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: null buffer");
    const double *a2s_bil = sa2s_bil->own, *a2s_cons = sa2s_cons->own, *o2s_bil = so2s_bil->own, *o2s_cons = so2s_cons->own;
    if (!as_bil || !as_cons || !os_bil || !os_cons) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: null table handle");
    if (!a2s_bil || !a2s_cons || !o2s_bil || !o2s_cons || !s2a || !s2o)
        return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: null buffer");
    if (members < 1) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: members must be >= 1");
    const int nS = as_bil->n_recv, nA = as_bil->n_send, nO = os_bil->n_send;
    if (as_cons->n_recv != nS || os_bil->n_recv != nS || os_cons->n_recv != nS || as_cons->n_send != nA ||
        os_cons->n_send != nO)
        return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: the four tables do not describe the same grid triple");
    SfcArgs a;
    a.as_bil = csr_of(as_bil); a.as_cons = csr_of(as_cons); a.os_bil = csr_of(os_bil); a.os_cons = csr_of(os_cons);
    a.a2s_bil = seg_of(sa2s_bil); a.a2s_cons = seg_of(sa2s_cons); a.o2s_bil = seg_of(so2s_bil); a.o2s_cons = seg_of(so2s_cons);
    a.s2a = s2a; a.s2o = s2o;
    a.has_full = full ? 1 : 0;
    if (full) a.full = *full; else memset(&a.full, 0, sizeof a.full);
    if (s_ld != 0 && s_ld < nS) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: s_ld < surface cells");
    if ((a_ld && a_ld < nA) || (o_ld && o_ld < nO)) return fail(DCCM_ERR_ARG, "dccm_sfc_exchange: a_ld / o_ld smaller than the tables' source extent");
    a.nA = a_ld ? a_ld : nA; a.nO = o_ld ? o_ld : nO; a.nS = nS; a.sld = s_ld ? s_ld : nS; a.M = members; a.sig1 = sig1;
    const int64_t n = (int64_t)nS * members;
    const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
    static const int minb = getenv("DCCM_SFC_MINB") ? atoi(getenv("DCCM_SFC_MINB")) : 5;   // tuning knob (5 measured best on B200, profiles/)
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (minb) {
    case 3: sfc_exchange_kernel<3><<<grid, kThreads, 0, st>>>(a); break;
    case 6: sfc_exchange_kernel<6><<<grid, kThreads, 0, st>>>(a); break;
    case 8: sfc_exchange_kernel<8><<<grid, kThreads, 0, st>>>(a); break;
    case 4: sfc_exchange_kernel<4><<<grid, kThreads, 0, st>>>(a); break;
    default: sfc_exchange_kernel<5><<<grid, kThreads, 0, st>>>(a); break;
    }
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}
