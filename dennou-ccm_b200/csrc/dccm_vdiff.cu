// dccm_vdiff.cu -- K3/K4: column-wise implicit vertical-diffusion / surface coupling solve.
//
// Replaces dcpam_sfc_implicit_coupling_mod (ref atm/dcpam_sfc_implicit_coupling_mod.f90):
//   K3 = SfcImplicitCoupling_VDiffForward (:72-378) + Solve_TriDiagSystem_Forward (:380-402)
//   K4 = SfcImplicitCoupling_VDiffBackward (:25-70) + Solve_TriDiagSystem_Backward (:404-418)
//
// One thread per column; arrays are (column, level) with the column index fastest, so every
// per-level access of a warp is one coalesced 256-byte request.  The reference builds three
// full tridiagonal matrices, copies two of them, and sweeps each system separately with
// whole-array statements (every 3-D array is streamed several times).  Here the transfer
// coefficients, the three matrix rows, the right-hand sides and the top-down elimination of
// one level are all evaluated in registers while the column is walked ONCE from k = kmax
// down to 2: each input is read exactly once and only what the backward pass needs is
// written -- the swept diagonal b'(k) of the three systems (a' == 1 and c' == 0 are implied)
// and the swept right-hand sides r'(k).
//
// Elimination stops at k = 2: the reference divides by Mtx(k=1,-1) = 0 at k = 1 (defect C-1,
// see DESIGN.md); row 1 is never read again, so RHS(1) is returned un-swept.
//
// FAST = false keeps the reference's operation order with IEEE divisions (the library is
// built with -fmad=false), which makes K3/K4 bit-identical to the reference loops.
// FAST = true shares reciprocals between the systems of one level (fewer fp64 divisions,
// <= a few ulp per level from the reference; verified to the 1e-12 tolerance).
#include <cuda_runtime.h>

#include <algorithm>

#include "dccm_arith.cuh"
#include "dccm_common.h"

using namespace dccm;

struct dccm_vdiff {
    int imax = 0, jmax = 0, kmax = 0, ncmax = 0, iq = 0;   // iq: 0-based water-vapour tracer
    double Grav = 0, CpDry = 0, GasRDry = 0, DelTime = 0;
    int fast = 0;
    int64_t NC = 0;
    int64_t coef_stride = 0;                                // 0 = NC
    double *bUV = nullptr, *bT = nullptr, *bQ = nullptr;   // swept diagonals, (NC, kmax)
    static constexpr int kRedoCap = 1 << 16;
    int *redo = nullptr;                                    // redo list of the reference-order forward solve (see FwdArgs)
    DevBuf in_buf, out_buf;                                 // scratch of the host entry points
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};     // H2D | kernel | D2H streams of the host entry points
};

namespace {

constexpr int kThreads = 128;
constexpr int kMaxTracer = 8;

struct FwdArgs {
    const double *FX, *FY, *FH, *FQ;
    const double *Press, *zExner, *rExner, *VirTemp, *Height, *DiffV, *DiffT, *DiffQ;
    double *DU, *DV, *DT, *DQ, *Coef1, *Coef2;
    double *bUV, *bT, *bQ;
    int64_t NC, cstride;          // cstride: slot stride of Coef1/Coef2 (>= NC)
    int64_t c0, c1;               // columns [c0, c1) of this launch (host entry points pipeline over column chunks)
    int K, iq;
    double Grav, CpDry, GasRDry, DelTime;
    int *redo;                    // [0] columns listed, [1] CTAs of the redo kernel done, [2] columns redone since creation
    int64_t *redo64;              // the listed columns
    int redo_cap;
};

// One column.  Arith supplies the divisions of the reference-order mode: FastArith (branch-free, same bits; the
// return value says whether every division stayed inside its fast path) or IeeeArith (plain operators).
template <int NQ, bool FAST, class Arith>
__device__ __forceinline__ bool forward_column(const FwdArgs &a, const int64_t c)
{
    const int64_t NC = a.NC;
    const int K = a.K;
    Arith ar;
    const double Grav = a.Grav, CpDry = a.CpDry, GasRDry = a.GasRDry;
    const double twodt = 2.0 * a.DelTime;
#define HL(p, l) __ldg(&(p)[c + NC * (int64_t)(l)])          /* half level l = 0..K */
#define FL(p, k) __ldg(&(p)[c + NC * (int64_t)((k) - 1)])    /* full level k = 1..K */
#define QH(n, l) __ldg(&a.FQ[c + NC * ((int64_t)(l) + (int64_t)(K + 1) * (n))])

    // state carried down the column (index "hi" = level k, "lo" = level k-1)
    double P_hi = HL(a.Press, K), H_hi = FL(a.Height, K);
    double zE_hi = FL(a.zExner, K), zE_hi2 = 0.0, rE_hi = 0.0;
    double TV_hi = 0.0, TT_hi = 0.0, TQ_hi = 0.0;              // transfer coefficients are 0 at k = kmax (:189-194)
    double FX_hi = HL(a.FX, K), FY_hi = HL(a.FY, K), FH_hi = HL(a.FH, K);
    double FQ_hi[NQ];
#pragma unroll
    for (int n = 0; n < NQ; n++) FQ_hi[n] = QH(n, K);
    double bUVn = 0.0, bTn = 0.0, bQn = 0.0, rUn = 0.0, rVn = 0.0, rTn = 0.0;
    double rQn[NQ];
#pragma unroll
    for (int n = 0; n < NQ; n++) rQn[n] = 0.0;
    double rU2 = 0.0, rV2 = 0.0, rT2 = 0.0, rQ2 = 0.0;
    // reference-order mode: refined reciprocals of the constant divisors, and of zExner carried down with it
    const double iGrav = FAST ? 0.0 : FastArith::prep(Grav), iTwodt = FAST ? 0.0 : FastArith::prep(twodt);
    double iz_hi = FAST ? 0.0 : FastArith::prep(zE_hi), iz_hi2 = 0.0, iz_lo = 0.0;

    for (int k = K; k >= 2; --k) {
        const int l = k - 1;
        const double P_lo = HL(a.Press, l), Tv = HL(a.VirTemp, l), H_lo = FL(a.Height, l);
        const double dV = HL(a.DiffV, l), dT = HL(a.DiffT, l), dQ = HL(a.DiffQ, l);
        const double rE_lo = HL(a.rExner, l), zE_lo = FL(a.zExner, l);
        const double FX_lo = HL(a.FX, l), FY_lo = HL(a.FY, l), FH_lo = HL(a.FH, l);
        double FQ_lo[NQ];
#pragma unroll
        for (int n = 0; n < NQ; n++) FQ_lo[n] = QH(n, l);

        double TV_lo, TT_lo, TQ_lo;
        double bUVp, bTp, bQp, rUp, rVp, rTp, rQp[NQ];
        const double rU = -(FX_hi - FX_lo), rV = -(FY_hi - FY_lo), rT = -(FH_hi - FH_lo);
        double rQ[NQ];
#pragma unroll
        for (int n = 0; n < NQ; n++) rQ[n] = -(FQ_hi[n] - FQ_lo[n]);
        if constexpr (FAST) {
            // transfer coefficients at half level l (:196-203)
            const double tmp = P_lo / (GasRDry * Tv * (H_hi - H_lo));
            TV_lo = dV * tmp; TT_lo = dT * tmp; TQ_lo = dQ * tmp;
            // matrix rows k (:207-293)
            const double mass = -(P_hi - P_lo) * (1.0 / (Grav * twodt));
            const double izhi = 1.0 / zE_hi;
            const double aT = -CpDry * rE_lo * TT_lo / zE_lo;
            double bT = CpDry * mass + CpDry * rE_lo * TT_lo * izhi;
            if (k < K) bT += CpDry * rE_hi * TT_hi * izhi;
            const double cT = (k < K) ? -CpDry * rE_hi * TT_hi / zE_hi2 : 0.0;
            const double aUV = -TV_lo, cUV = -TV_hi;
            double bUV = mass + TV_lo;
            if (k < K) bUV = bUV + TV_hi;
            const double aQ = -TQ_lo, cQ = -TQ_hi;
            double bQ = mass + TQ_lo;
            if (k < K) bQ = bQ + TQ_hi;
            // top-down elimination of level k (:388-400)
            if (k == K) {
                const double iu = 1.0 / aUV, it = 1.0 / aT, iq = 1.0 / aQ;
                bUVp = bUV * iu; rUp = rU * iu; rVp = rV * iu;
                bTp = bT * it; rTp = rT * it;
                bQp = bQ * iq;
#pragma unroll
                for (int n = 0; n < NQ; n++) rQp[n] = rQ[n] * iq;
            } else {
                const double dUV = aUV * bUVn, dT_ = aT * bTn, dQ_ = aQ * bQn;
                const double iu = 1.0 / dUV, it = 1.0 / dT_, iq = 1.0 / dQ_;
                bUVp = (bUV * bUVn - cUV) * iu;
                rUp = (rU * bUVn - cUV * rUn) * iu;
                rVp = (rV * bUVn - cUV * rVn) * iu;
                bTp = (bT * bTn - cT) * it;
                rTp = (rT * bTn - cT * rTn) * it;
                bQp = (bQ * bQn - cQ) * iq;
#pragma unroll
                for (int n = 0; n < NQ; n++) rQp[n] = (rQ[n] * bQn - cQ * rQn[n]) * iq;
            }
        } else {
            // Reference order, every quotient with the bits of the IEEE operator.  The ~17 divisions of a level have
            // only 8 distinct divisors (two of them constants, two carried down from the level above): the refined
            // reciprocal of the division sequence is computed once per divisor (FastArith::prep / div_by) and the
            // per-division range test is accumulated instead of branched on; a column whose test failed anywhere
            // (zero / denormal / non-finite operands) goes to the redo list and is solved again with the plain
            // operators by vdiff_forward_redo_kernel.
            iz_lo = FastArith::prep(zE_lo);
            {
                // transfer coefficients at half level l (:196-203)
                const double gt = GasRDry * Tv, dH = H_hi - H_lo;
                const double tmp = ar.div_by(ar.div_by(P_lo, gt, ar.prep(gt)), dH, ar.prep(dH));
                TV_lo = dV * tmp; TT_lo = dT * tmp; TQ_lo = dQ * tmp;
                // matrix rows k (:207-293)
                const double dP = P_hi - P_lo;
                const double mass = ar.div_by(ar.div_by(-dP, Grav, iGrav), twodt, iTwodt);
                const double aT = ar.div_by(-CpDry * rE_lo, zE_lo, iz_lo) * TT_lo;
                double bT = ar.div_by(ar.div_by(-CpDry * dP, Grav, iGrav), twodt, iTwodt)
                          + ar.div_by(CpDry * rE_lo, zE_hi, iz_hi) * TT_lo;
                if (k < K) bT = bT + ar.div_by(CpDry * rE_hi, zE_hi, iz_hi) * TT_hi;
                const double cT = (k < K) ? ar.div_by(-CpDry * rE_hi, zE_hi2, iz_hi2) * TT_hi : 0.0;
                const double aUV = -TV_lo, cUV = -TV_hi;
                double bUV = mass + TV_lo;
                if (k < K) bUV = bUV + TV_hi;
                const double aQ = -TQ_lo, cQ = -TQ_hi;
                double bQ = mass + TQ_lo;
                if (k < K) bQ = bQ + TQ_hi;
                // top-down elimination of level k (:388-400)
                if (k == K) {
                    const double tu = ar.prep(aUV), tt = ar.prep(aT), tq = ar.prep(aQ);
                    bUVp = ar.div_by(bUV, aUV, tu); rUp = ar.div_by(rU, aUV, tu); rVp = ar.div_by(rV, aUV, tu);
                    bTp = ar.div_by(bT, aT, tt); rTp = ar.div_by(rT, aT, tt);
                    bQp = ar.div_by(bQ, aQ, tq);
#pragma unroll
                    for (int n = 0; n < NQ; n++) rQp[n] = ar.div_by(rQ[n], aQ, tq);
                } else {
                    const double dUV = aUV * bUVn, dT_ = aT * bTn, dQ_ = aQ * bQn;
                    const double tu = ar.prep(dUV), tt = ar.prep(dT_), tq = ar.prep(dQ_);
                    bUVp = ar.div_by(bUV * bUVn - cUV, dUV, tu);
                    rUp = ar.div_by(rU * bUVn - cUV * rUn, dUV, tu);
                    rVp = ar.div_by(rV * bUVn - cUV * rVn, dUV, tu);
                    bTp = ar.div_by(bT * bTn - cT, dT_, tt);
                    rTp = ar.div_by(rT * bTn - cT * rTn, dT_, tt);
                    bQp = ar.div_by(bQ * bQn - cQ, dQ_, tq);
#pragma unroll
                    for (int n = 0; n < NQ; n++) rQp[n] = ar.div_by(rQ[n] * bQn - cQ * rQn[n], dQ_, tq);
                }
            }
        }
        const int64_t o = c + NC * (int64_t)(k - 1);
        a.bUV[o] = bUVp; a.bT[o] = bTp; a.bQ[o] = bQp;
        a.DU[o] = rUp; a.DV[o] = rVp; a.DT[o] = rTp;
#pragma unroll
        for (int n = 0; n < NQ; n++) a.DQ[o + NC * (int64_t)K * n] = rQp[n];

        // shift down one level
        P_hi = P_lo; H_hi = H_lo; zE_hi2 = zE_hi; zE_hi = zE_lo; rE_hi = rE_lo;
        iz_hi2 = iz_hi; iz_hi = iz_lo;
        TV_hi = TV_lo; TT_hi = TT_lo; TQ_hi = TQ_lo;
        FX_hi = FX_lo; FY_hi = FY_lo; FH_hi = FH_lo;
        bUVn = bUVp; bTn = bTp; bQn = bQp; rUn = rUp; rVn = rVp; rTn = rTp;
#pragma unroll
        for (int n = 0; n < NQ; n++) { FQ_hi[n] = FQ_lo[n]; rQn[n] = rQp[n]; }
        if (k == 2) {
            rU2 = rUp; rV2 = rVp; rT2 = rTp;
#pragma unroll
            for (int n = 0; n < NQ; n++) if (n == a.iq) rQ2 = rQp[n];
        }
    }

    // level 1: un-swept RHS (= Coef2 before the correction, :313-316) and the coupling
    // coefficients (:344-376).  Now "hi" = level 1.
    {
        const double P0 = HL(a.Press, 0);
        const double rU1 = -(FX_hi - HL(a.FX, 0)), rV1 = -(FY_hi - HL(a.FY, 0)), rT1 = -(FH_hi - HL(a.FH, 0));
        double rQ1v = 0.0;
#pragma unroll
        for (int n = 0; n < NQ; n++) {
            const double r = -(FQ_hi[n] - QH(n, 0));
            a.DQ[c + NC * (int64_t)K * n] = r;
            if (n == a.iq) rQ1v = r;
        }
        a.DU[c] = rU1; a.DV[c] = rV1; a.DT[c] = rT1;

        const double tmp1 = -(P_hi - P0) / Grav / twodt;
        const double DFADUV1 = TV_hi, DFADUV2 = -TV_hi;
        const double c1uv = tmp1 + DFADUV1 - DFADUV2 / bUVn;
        a.Coef1[c] = c1uv;
        a.Coef2[c] = rU1 - DFADUV2 * rU2 / bUVn;
        const int64_t CS = a.cstride;
        a.Coef1[c + CS] = c1uv;
        a.Coef2[c + CS] = rV1 - DFADUV2 * rV2 / bUVn;
        const double DFADT1 = CpDry * rE_hi * TT_hi / zE_hi;
        const double DFADT2 = -CpDry * rE_hi * TT_hi / zE_hi2;
        a.Coef1[c + 2 * CS] = CpDry * tmp1 + DFADT1 - DFADT2 / bTn;
        a.Coef2[c + 2 * CS] = rT1 - DFADT2 * rT2 / bTn;
        const double DFADQ1 = TQ_hi, DFADQ2 = -TQ_hi;
        a.Coef1[c + 3 * CS] = tmp1 + DFADQ1 - DFADQ2 / bQn;
        a.Coef2[c + 3 * CS] = rQ1v - DFADQ2 * rQ2 / bQn;
    }
#undef HL
#undef FL
#undef QH
    return ar.good();
}

// 5 CTAs (20 warps) per SM for one or two tracers -- more loads in flight, 4.83 -> 4.61 ms at config 5
template <int NQ, bool FAST, int MINB = (NQ <= 2 ? 5 : 1)>
__global__ void __launch_bounds__(kThreads, MINB) vdiff_forward_kernel(const FwdArgs a)
{
    const int64_t c = a.c0 + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= a.c1) return;
    if (!forward_column<NQ, FAST, FastArith>(a, c)) {
        const int slot = atomicAdd(&a.redo[0], 1);
        if (slot < a.redo_cap) a.redo64[slot] = c;
    }
}

// The columns the branch-free divisions did not accept (normally none: an empty launch), solved again with the plain
// IEEE operators; more columns than the list holds: every column of the launch is redone.  The last CTA clears the list.
template <int NQ>
__global__ void __launch_bounds__(kThreads) vdiff_forward_redo_kernel(const FwdArgs a)
{
    const int listed = a.redo[0];
    if (listed <= 0) return;
    const int64_t stride = (int64_t)gridDim.x * kThreads, t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (listed <= a.redo_cap) {
        for (int64_t t = t0; t < listed; t += stride) forward_column<NQ, false, IeeeArith>(a, a.redo64[t]);
    } else {
        for (int64_t c = a.c0 + t0; c < a.c1; c += stride) forward_column<NQ, false, IeeeArith>(a, c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&a.redo[1], 1) == (int)gridDim.x - 1) { a.redo[2] += listed; a.redo[0] = 0; a.redo[1] = 0; __threadfence(); }
    }
}

struct BwdArgs {
    double *DU, *DV, *DT, *DQ;
    const double *bUV, *bT, *bQ;
    const double *level1;
    int64_t NC, c0, c1;           // columns [c0, c1) of this launch
    int K, iq;
    double DelTime;
};

template <int NQ>
__global__ void __launch_bounds__(kThreads) vdiff_backward_kernel(const BwdArgs a)
{
    const int64_t c = a.c0 + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t NC = a.NC;
    if (c >= a.c1) return;
    const int K = a.K;
    const double twodt = 2.0 * a.DelTime;
    // level 1: the surface-layer increment delivered by the surface component
    // (ref atm/dccm_atm_mod.f90:832-835)
    double xU, xV, xT, xQ[NQ];
    if (a.level1) {
        xU = a.level1[c]; xV = a.level1[c + NC]; xT = a.level1[c + 2 * NC];
    } else {
        xU = a.DU[c]; xV = a.DV[c]; xT = a.DT[c];
    }
#pragma unroll
    for (int n = 0; n < NQ; n++) {
        xQ[n] = a.DQ[c + NC * (int64_t)K * n];
        if (a.level1 && n == a.iq) xQ[n] = a.level1[c + 3 * NC];
    }
    a.DU[c] = xU / twodt; a.DV[c] = xV / twodt; a.DT[c] = xT / twodt;           // :54-63
#pragma unroll
    for (int n = 0; n < NQ; n++) a.DQ[c + NC * (int64_t)K * n] = xQ[n] / twodt;
    // Bottom-up substitution (:414-416).  The recurrence x_k = (r'_k - x_{k-1}) / b'_k is a serial
    // chain of divisions, but its operands are not: the loads of level k+2 are issued before level k
    // is consumed (two register sets, loop unrolled by two), so memory latency overlaps the chain.
    struct Lvl { double bUV, bT, bQ, rU, rV, rT, rQ[NQ]; };
    auto load = [&](int k, Lvl &v) {
        const int64_t o = c + NC * (int64_t)(k - 1);
        v.bUV = __ldg(&a.bUV[o]); v.bT = __ldg(&a.bT[o]); v.bQ = __ldg(&a.bQ[o]);
        v.rU = a.DU[o]; v.rV = a.DV[o]; v.rT = a.DT[o];
#pragma unroll
        for (int n = 0; n < NQ; n++) v.rQ[n] = a.DQ[o + NC * (int64_t)K * n];
    };
    auto solve = [&](int k, const Lvl &v) {
        const int64_t o = c + NC * (int64_t)(k - 1);
        xU = (v.rU - xU) / v.bUV;
        xV = (v.rV - xV) / v.bUV;
        xT = (v.rT - xT) / v.bT;
        a.DU[o] = xU / twodt; a.DV[o] = xV / twodt; a.DT[o] = xT / twodt;
#pragma unroll
        for (int n = 0; n < NQ; n++) {
            xQ[n] = (v.rQ[n] - xQ[n]) / v.bQ;
            a.DQ[o + NC * (int64_t)K * n] = xQ[n] / twodt;
        }
    };
    Lvl s0, s1;
    load(2, s0);
    if (K >= 3) load(3, s1);
    for (int k = 2; k <= K; k += 2) {
        const Lvl c0 = s0;
        if (k + 2 <= K) load(k + 2, s0);
        solve(k, c0);
        if (k + 1 <= K) {
            const Lvl c1 = s1;
            if (k + 3 <= K) load(k + 3, s1);
            solve(k + 1, c1);
        }
    }
}

void launch_forward_redo(int nq, const FwdArgs &a, unsigned grid, cudaStream_t st)
{
    switch (nq) {
    case 1: vdiff_forward_redo_kernel<1><<<grid, kThreads, 0, st>>>(a); break;
    case 2: vdiff_forward_redo_kernel<2><<<grid, kThreads, 0, st>>>(a); break;
    case 3: vdiff_forward_redo_kernel<3><<<grid, kThreads, 0, st>>>(a); break;
    case 4: vdiff_forward_redo_kernel<4><<<grid, kThreads, 0, st>>>(a); break;
    case 5: vdiff_forward_redo_kernel<5><<<grid, kThreads, 0, st>>>(a); break;
    case 6: vdiff_forward_redo_kernel<6><<<grid, kThreads, 0, st>>>(a); break;
    case 7: vdiff_forward_redo_kernel<7><<<grid, kThreads, 0, st>>>(a); break;
    default: vdiff_forward_redo_kernel<8><<<grid, kThreads, 0, st>>>(a); break;
    }
}

template <bool FAST>
void launch_forward(int nq, const FwdArgs &a, unsigned grid, cudaStream_t st)
{
    switch (nq) {
    case 1: vdiff_forward_kernel<1, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 2: vdiff_forward_kernel<2, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 3: vdiff_forward_kernel<3, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 4: vdiff_forward_kernel<4, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 5: vdiff_forward_kernel<5, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 6: vdiff_forward_kernel<6, FAST><<<grid, kThreads, 0, st>>>(a); break;
    case 7: vdiff_forward_kernel<7, FAST><<<grid, kThreads, 0, st>>>(a); break;
    default: vdiff_forward_kernel<8, FAST><<<grid, kThreads, 0, st>>>(a); break;
    }
}

}  // namespace

extern "C" int dccm_vdiff_create(int imax, int jmax, int kmax, int ncmax, int index_h2ovap,
                                 double Grav, double CpDry, double GasRDry, double DelTime, dccm_vdiff **out)
{
    *out = nullptr;
    if (imax < 1 || jmax < 1) return fail(DCCM_ERR_ARG, "dccm_vdiff_create: bad horizontal size");
    if (kmax < 2) return fail(DCCM_ERR_ARG, "dccm_vdiff_create: kmax must be >= 2 (the coupling coefficients use level 2)");
    if (ncmax < 1 || ncmax > kMaxTracer)
        return fail(DCCM_ERR_ARG, "dccm_vdiff_create: ncmax=%d outside 1..%d", ncmax, kMaxTracer);
    if (index_h2ovap < 1 || index_h2ovap > ncmax) return fail(DCCM_ERR_ARG, "dccm_vdiff_create: IndexH2OVap out of range");
    if (!(DelTime > 0.0) || !(Grav > 0.0)) return fail(DCCM_ERR_ARG, "dccm_vdiff_create: DelTime and Grav must be positive");
    int rc = ensure_device();
    if (rc) return rc;
    dccm_vdiff *h = new dccm_vdiff();
    h->imax = imax; h->jmax = jmax; h->kmax = kmax; h->ncmax = ncmax; h->iq = index_h2ovap - 1;
    h->Grav = Grav; h->CpDry = CpDry; h->GasRDry = GasRDry; h->DelTime = DelTime;
    h->NC = (int64_t)imax * jmax;
    size_t bytes = sizeof(double) * (size_t)h->NC * kmax;
    cudaError_t e = cudaMalloc(&h->bUV, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->bT, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->bQ, bytes);
    const size_t redo_bytes = 16 + sizeof(int64_t) * dccm_vdiff::kRedoCap;
    if (e == cudaSuccess) e = cudaMalloc(&h->redo, redo_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->redo, 0, redo_bytes);
    if (e != cudaSuccess) {
        dccm_vdiff_destroy(h);
        return fail(DCCM_ERR_CUDA, "dccm_vdiff_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return DCCM_OK;
}

extern "C" void dccm_vdiff_destroy(dccm_vdiff *h)
{
    if (!h) return;
    cudaFree(h->bUV); cudaFree(h->bT); cudaFree(h->bQ); cudaFree(h->redo);
    h->in_buf.release(); h->out_buf.release();
    for (cudaStream_t st : h->pipe) if (st) cudaStreamDestroy(st);
    delete h;
}

extern "C" int dccm_vdiff_set_mode(dccm_vdiff *h, int fast)
{
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_set_mode: null handle");
    h->fast = fast ? 1 : 0;
    return DCCM_OK;
}

extern "C" int dccm_vdiff_redo_total(dccm_vdiff *h, int64_t *columns)
{
    if (!h || !columns) return fail(DCCM_ERR_ARG, "dccm_vdiff_redo_total: null argument");
    int v = 0;
    DCCM_CUDA_TRY(cudaDeviceSynchronize());
    DCCM_CUDA_TRY(cudaMemcpy(&v, h->redo + 2, sizeof v, cudaMemcpyDeviceToHost));
    *columns = v;
    return DCCM_OK;
}

extern "C" int dccm_vdiff_set_coef_stride(dccm_vdiff *h, int64_t slot_stride)
{
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_set_coef_stride: null handle");
    if (slot_stride != 0 && slot_stride < h->NC) return fail(DCCM_ERR_ARG, "dccm_vdiff_set_coef_stride: stride < columns");
    h->coef_stride = slot_stride;
    return DCCM_OK;
}

namespace {
int forward_range(dccm_vdiff *h,
    const double *FX, const double *FY, const double *FH, const double *FQ,
    const double *Press, const double *zExner, const double *rExner,
    const double *VirTemp, const double *Height,
    const double *DiffV, const double *DiffT, const double *DiffQ,
    double *DU, double *DV, double *DT, double *DQ, double *Coef1, double *Coef2,
    int64_t c0, int64_t c1, cudaStream_t st)
{
    FwdArgs a;
    a.FX = FX; a.FY = FY; a.FH = FH; a.FQ = FQ; a.Press = Press; a.zExner = zExner; a.rExner = rExner;
    a.VirTemp = VirTemp; a.Height = Height; a.DiffV = DiffV; a.DiffT = DiffT; a.DiffQ = DiffQ;
    a.DU = DU; a.DV = DV; a.DT = DT; a.DQ = DQ; a.Coef1 = Coef1; a.Coef2 = Coef2;
    a.bUV = h->bUV; a.bT = h->bT; a.bQ = h->bQ;
    a.NC = h->NC; a.K = h->kmax; a.iq = h->iq; a.c0 = c0; a.c1 = c1;
    a.cstride = h->coef_stride > 0 ? h->coef_stride : h->NC;
    a.Grav = h->Grav; a.CpDry = h->CpDry; a.GasRDry = h->GasRDry; a.DelTime = h->DelTime;
    a.redo = h->redo; a.redo64 = reinterpret_cast<int64_t *>(h->redo + 4); a.redo_cap = dccm_vdiff::kRedoCap;
    const unsigned grid = (unsigned)((c1 - c0 + kThreads - 1) / kThreads);
    if (h->fast) launch_forward<true>(h->ncmax, a, grid, st);
    else {
        launch_forward<false>(h->ncmax, a, grid, st);
        DCCM_CUDA_TRY(cudaGetLastError());
        launch_forward_redo(h->ncmax, a, std::min(grid, 4u * (unsigned)num_sms()), st);
    }
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

int backward_range(dccm_vdiff *h, double *DU, double *DV, double *DT, double *DQ, const double *level1,
                   int64_t c0, int64_t c1, cudaStream_t st)
{
    BwdArgs a;
    a.DU = DU; a.DV = DV; a.DT = DT; a.DQ = DQ;
    a.bUV = h->bUV; a.bT = h->bT; a.bQ = h->bQ; a.level1 = level1;
    a.NC = h->NC; a.K = h->kmax; a.iq = h->iq; a.DelTime = h->DelTime; a.c0 = c0; a.c1 = c1;
    const unsigned grid = (unsigned)((c1 - c0 + kThreads - 1) / kThreads);
    switch (h->ncmax) {
    case 1: vdiff_backward_kernel<1><<<grid, kThreads, 0, st>>>(a); break;
    case 2: vdiff_backward_kernel<2><<<grid, kThreads, 0, st>>>(a); break;
    case 3: vdiff_backward_kernel<3><<<grid, kThreads, 0, st>>>(a); break;
    case 4: vdiff_backward_kernel<4><<<grid, kThreads, 0, st>>>(a); break;
    case 5: vdiff_backward_kernel<5><<<grid, kThreads, 0, st>>>(a); break;
    case 6: vdiff_backward_kernel<6><<<grid, kThreads, 0, st>>>(a); break;
    case 7: vdiff_backward_kernel<7><<<grid, kThreads, 0, st>>>(a); break;
    default: vdiff_backward_kernel<8><<<grid, kThreads, 0, st>>>(a); break;
    }
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

// Host entry points move their arguments in column chunks: chunk j+1 is on its way in (H2D stream) while the
// kernel runs on chunk j (kernel stream) and chunk j-1 is on its way out (D2H stream); arrays keep the
// reference layout (column fastest, level slowest), so a chunk of an array is a 2-D copy of `levels` rows.
// PCIe is full duplex: the outputs ride for free under the inputs.  Pinned host arrays overlap; pageable ones
// are staged by the driver and simply serialise.
struct Pipe {
    dccm_vdiff *h;
    std::vector<cudaEvent_t> ev;
    int nchunk;
    int64_t NC;
    explicit Pipe(dccm_vdiff *hh) : h(hh), NC(hh->NC)
    {
        const char *env = getenv("DCCM_HOST_CHUNKS");         // testing knob: force a chunk count on small problems
        nchunk = env ? std::max(1, atoi(env)) : (NC >= (1 << 19) ? 12 : 1);
        if ((int64_t)nchunk > (NC + 31) / 32) nchunk = (int)((NC + 31) / 32);
    }
    int init()
    {
        for (auto &st : h->pipe)
            if (!st) DCCM_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        ev.resize(2 * (size_t)nchunk);
        for (auto &e : ev) DCCM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return DCCM_OK;
    }
    ~Pipe() { for (auto e : ev) if (e) cudaEventDestroy(e); }
    void range(int j, int64_t &c0, int64_t &c1) const
    {
        const int64_t per = ((NC + nchunk - 1) / nchunk + 31) / 32 * 32;
        c0 = std::min<int64_t>(NC, per * j); c1 = std::min<int64_t>(NC, per * (j + 1));
    }
    cudaError_t copy(double *dst, const double *src, int64_t c0, int64_t c1, size_t rows, cudaMemcpyKind kind, cudaStream_t st) const
    {
        return cudaMemcpy2DAsync(dst + c0, sizeof(double) * NC, src + c0, sizeof(double) * NC,
                                 sizeof(double) * (size_t)(c1 - c0), rows, kind, st);
    }
};
}  // namespace

extern "C" int dccm_vdiff_forward_device(dccm_vdiff *h,
    const double *FX, const double *FY, const double *FH, const double *FQ,
    const double *Press, const double *zExner, const double *rExner,
    const double *VirTemp, const double *Height,
    const double *DiffV, const double *DiffT, const double *DiffQ,
    double *DU, double *DV, double *DT, double *DQ, double *Coef1, double *Coef2, void *stream)
{
    NvtxRange nvtx("dccm_vdiff_forward_device");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_forward: null handle");
    return forward_range(h, FX, FY, FH, FQ, Press, zExner, rExner, VirTemp, Height, DiffV, DiffT, DiffQ,
                         DU, DV, DT, DQ, Coef1, Coef2, 0, h->NC, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dccm_vdiff_backward_device(dccm_vdiff *h, double *DU, double *DV, double *DT, double *DQ,
                                          const double *level1, void *stream)
{
    NvtxRange nvtx("dccm_vdiff_backward_device");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_backward: null handle");
    return backward_range(h, DU, DV, DT, DQ, level1, 0, h->NC, reinterpret_cast<cudaStream_t>(stream));
}

// columns [c0, c1) only: the pieces of a latitude-slab pipeline (exchange.SurfaceExchange.step_pipelined)
extern "C" int dccm_vdiff_forward_cols_device(dccm_vdiff *h,
    const double *FX, const double *FY, const double *FH, const double *FQ,
    const double *Press, const double *zExner, const double *rExner,
    const double *VirTemp, const double *Height,
    const double *DiffV, const double *DiffT, const double *DiffQ,
    double *DU, double *DV, double *DT, double *DQ, double *Coef1, double *Coef2,
    int64_t c0, int64_t c1, void *stream)
{
    NvtxRange nvtx("dccm_vdiff_forward_cols_device");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_forward: null handle");
    if (c0 < 0 || c1 > h->NC || c0 >= c1) return fail(DCCM_ERR_ARG, "dccm_vdiff_forward_cols: bad column range");
    return forward_range(h, FX, FY, FH, FQ, Press, zExner, rExner, VirTemp, Height, DiffV, DiffT, DiffQ,
                         DU, DV, DT, DQ, Coef1, Coef2, c0, c1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dccm_vdiff_backward_cols_device(dccm_vdiff *h, double *DU, double *DV, double *DT, double *DQ,
                                               const double *level1, int64_t c0, int64_t c1, void *stream)
{
    NvtxRange nvtx("dccm_vdiff_backward_cols_device");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_backward: null handle");
    if (c0 < 0 || c1 > h->NC || c0 >= c1) return fail(DCCM_ERR_ARG, "dccm_vdiff_backward_cols: bad column range");
    return backward_range(h, DU, DV, DT, DQ, level1, c0, c1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dccm_vdiff_forward_host(dccm_vdiff *h,
    const double *FX, const double *FY, const double *FH, const double *FQ,
    const double *Press, const double *zExner, const double *rExner,
    const double *VirTemp, const double *Height,
    const double *DiffV, const double *DiffT, const double *DiffQ,
    double *DU, double *DV, double *DT, double *DQ, double *Coef1, double *Coef2)
{
    NvtxRange nvtx("dccm_vdiff_forward_host");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_forward: null handle");
    const size_t NC = (size_t)h->NC, K = h->kmax, nc = h->ncmax;
    const size_t half = NC * (K + 1), full = NC * K;
    // inputs: 9 half-level + 2 full-level + nc half-level tracer fluxes
    int rc = h->in_buf.reserve(sizeof(double) * (half * (9 + nc) + full * 2));
    if (rc) return rc;
    rc = h->out_buf.reserve(sizeof(double) * (full * (3 + nc) + NC * 8));
    if (rc) return rc;
    double *p = h->in_buf.as<double>();
    auto take = [&](size_t n) { double *q = p; p += n; return q; };
    double *dFX = take(half), *dFY = take(half), *dFH = take(half), *dFQ = take(half * nc);
    double *dP = take(half), *dzE = take(full), *drE = take(half), *dTv = take(half), *dH = take(full);
    double *dDV = take(half), *dDT = take(half), *dDQ = take(half);
    p = h->out_buf.as<double>();
    double *oDU = take(full), *oDV = take(full), *oDT = take(full), *oDQ = take(full * nc);
    double *oC1 = take(NC * 4), *oC2 = take(NC * 4);
    if (h->coef_stride > 0 && h->coef_stride != h->NC)
        return fail(DCCM_ERR_ARG, "dccm_vdiff_forward_host: a coefficient slot stride is only meaningful for device buffers");
    Pipe pp(h);
    rc = pp.init();
    if (rc) return rc;
    cudaStream_t sin = h->pipe[0], sk = h->pipe[1], sout = h->pipe[2];
    const size_t Kh = K + 1;
    for (int j = 0; j < pp.nchunk; j++) {
        int64_t c0, c1;
        pp.range(j, c0, c1);
        if (c0 >= c1) continue;
        auto in = [&](double *d, const double *s, size_t rows) { return pp.copy(d, s, c0, c1, rows, cudaMemcpyHostToDevice, sin); };
        DCCM_CUDA_TRY(in(dFX, FX, Kh)); DCCM_CUDA_TRY(in(dFY, FY, Kh)); DCCM_CUDA_TRY(in(dFH, FH, Kh));
        DCCM_CUDA_TRY(in(dFQ, FQ, Kh * nc)); DCCM_CUDA_TRY(in(dP, Press, Kh)); DCCM_CUDA_TRY(in(dzE, zExner, K));
        DCCM_CUDA_TRY(in(drE, rExner, Kh)); DCCM_CUDA_TRY(in(dTv, VirTemp, Kh)); DCCM_CUDA_TRY(in(dH, Height, K));
        DCCM_CUDA_TRY(in(dDV, DiffV, Kh)); DCCM_CUDA_TRY(in(dDT, DiffT, Kh)); DCCM_CUDA_TRY(in(dDQ, DiffQ, Kh));
        DCCM_CUDA_TRY(cudaEventRecord(pp.ev[2 * j], sin));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sk, pp.ev[2 * j], 0));
        rc = forward_range(h, dFX, dFY, dFH, dFQ, dP, dzE, drE, dTv, dH, dDV, dDT, dDQ, oDU, oDV, oDT, oDQ, oC1, oC2, c0, c1, sk);
        if (rc) return rc;
        DCCM_CUDA_TRY(cudaEventRecord(pp.ev[2 * j + 1], sk));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sout, pp.ev[2 * j + 1], 0));
        auto out = [&](double *d, const double *s, size_t rows) { return pp.copy(d, s, c0, c1, rows, cudaMemcpyDeviceToHost, sout); };
        DCCM_CUDA_TRY(out(DU, oDU, K)); DCCM_CUDA_TRY(out(DV, oDV, K)); DCCM_CUDA_TRY(out(DT, oDT, K));
        DCCM_CUDA_TRY(out(DQ, oDQ, K * nc)); DCCM_CUDA_TRY(out(Coef1, oC1, 4)); DCCM_CUDA_TRY(out(Coef2, oC2, 4));
    }
    DCCM_CUDA_TRY(cudaStreamSynchronize(sout));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sk));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sin));
    return DCCM_OK;
}

extern "C" int dccm_vdiff_backward_host(dccm_vdiff *h, double *DU, double *DV, double *DT, double *DQ)
{
    NvtxRange nvtx("dccm_vdiff_backward_host");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_vdiff_backward: null handle");
    const size_t NC = (size_t)h->NC, K = h->kmax, nc = h->ncmax, full = NC * K;
    int rc = h->out_buf.reserve(sizeof(double) * (full * (3 + nc) + NC * 8));
    if (rc) return rc;
    double *p = h->out_buf.as<double>();
    double *oDU = p, *oDV = p + full, *oDT = p + 2 * full, *oDQ = p + 3 * full;
    Pipe pp(h);
    rc = pp.init();
    if (rc) return rc;
    cudaStream_t sin = h->pipe[0], sk = h->pipe[1], sout = h->pipe[2];
    for (int j = 0; j < pp.nchunk; j++) {
        int64_t c0, c1;
        pp.range(j, c0, c1);
        if (c0 >= c1) continue;
        DCCM_CUDA_TRY(pp.copy(oDU, DU, c0, c1, K, cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(pp.copy(oDV, DV, c0, c1, K, cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(pp.copy(oDT, DT, c0, c1, K, cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(pp.copy(oDQ, DQ, c0, c1, K * nc, cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(cudaEventRecord(pp.ev[2 * j], sin));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sk, pp.ev[2 * j], 0));
        rc = backward_range(h, oDU, oDV, oDT, oDQ, nullptr, c0, c1, sk);
        if (rc) return rc;
        DCCM_CUDA_TRY(cudaEventRecord(pp.ev[2 * j + 1], sk));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sout, pp.ev[2 * j + 1], 0));
        DCCM_CUDA_TRY(pp.copy(DU, oDU, c0, c1, K, cudaMemcpyDeviceToHost, sout));
        DCCM_CUDA_TRY(pp.copy(DV, oDV, c0, c1, K, cudaMemcpyDeviceToHost, sout));
        DCCM_CUDA_TRY(pp.copy(DT, oDT, c0, c1, K, cudaMemcpyDeviceToHost, sout));
        DCCM_CUDA_TRY(pp.copy(DQ, oDQ, c0, c1, K * nc, cudaMemcpyDeviceToHost, sout));
    }
    DCCM_CUDA_TRY(cudaStreamSynchronize(sout));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sk));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sin));
    return DCCM_OK;
}
