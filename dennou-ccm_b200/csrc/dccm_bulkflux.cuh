// dccm_bulkflux.cuh -- per-column bulk surface flux + implicit surface-layer update, in
// registers.  Shared by the stand-alone kernel (dccm_bulkflux.cu) and the fused
// remap -> bulk-flux kernel (dccm_exchange.cu).
//
// Restates DSFCM_Util_SfcBulkFlux_Get + BulkCoefL82 for ONE column
// (ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:194-415, :478-568), keeping the reference's
// operation order (the library is compiled with -fmad=false); exp / log / ** are the fixed IEEE
// sequences of dccm_pmath.cuh (within 1 ulp of any libm, same bits on every machine), so the whole
// column is reproducible bit for bit by a CPU that follows the same definitions.
// The reference's ~25 full-grid temporaries (:157-186) and six OpenMP passes collapse into
// registers: each input is read once and each output written once.
#pragma once
#include "dccm_arith.cuh"
#include "dccm_pmath.cuh"

namespace dccm {

// ref sfc/DSFCM_Util_SfcBulkFlux_mod.f90:35-58 (note: these are the SFC component's own
// constants -- CpDry = 1616, MolWt 0.018 -- not DCPAM's; reproduced, not reconciled)
namespace sfc {
constexpr double FKarm = 0.4;
constexpr double GasRUniv = 8.3144621;
constexpr double StB = 5.670373e-8;
constexpr double Grav = 9.8;
constexpr double MolWtDry = 1.8e-2;
constexpr double MolWtWet = 1.8e-2;
constexpr double CpDry = 1616.0;
constexpr double LatentHeat = 2425300.0;
constexpr double LatentHeatFusion = 334000.0;
constexpr double RefPress = 1e5;
constexpr double GasRDry = GasRUniv / MolWtDry;
constexpr double GasRWet = GasRUniv / MolWtWet;
constexpr double EpsV = MolWtWet / MolWtDry;
constexpr double Es0 = 611.0;
constexpr double HumidCoef = 1.0;
constexpr double RoughLength = 1e-4;
constexpr double RoughLenHeatFactor = 1.0;
// limits, ref :574-591 (read_config: hard-coded, not namelist)
constexpr double VelMinForRi = 0.01, VelMinForVel = 0.01, VelMinForTemp = 0.01, VelMinForQVap = 0.01;
constexpr double VelMaxForVel = 1000.0, VelMaxForTemp = 1000.0, VelMaxForQVap = 1000.0;
constexpr double VelBulkCoefMin = 0.0, TempBulkCoefMin = 0.0, QVapBulkCoefMin = 0.0;
constexpr double VelBulkCoefMax = 1.0, TempBulkCoefMax = 1.0, QVapBulkCoefMax = 1.0;
}  // namespace sfc

struct BulkIn {
    double WindU, WindV, SfcAirTemp, QVap1, SDwRFlx, LDwRFlx;
    double Coef1[4], Coef2[4];
    double SfcTemp[2], SfcAlbedo[2];
    double SIceCon, SfcHeight, SfcPress;
};

struct BulkOut {
    double WindStressX[3], WindStressY[3], SenHFlx[3], QVapMFlx[3], LatHFlx[3];
    double VelTC[3], TempTC[3], QVapTC[3];
    double Del[4];
    double SUwRFlx[3], LUwRFlx[3];
    double HFlx_ns[2], HFlx_sr[2], DHFlxDTs[2];
    double SfcTemp3, SfcAlbedo3;
};

// What the implicit update (phase 2) needs from the flux evaluation (phase 1) besides BulkOut.
struct BulkMid {
    double Exner, SfcExner, iEx;            // iEx: refined reciprocal of Exner (Arith::prep)
    double Frac[2], QVapSat[2];
};

// Phase 1 (:194-349): per-slot bulk coefficients, transfer coefficients, fluxes and their
// area-weighted composite.  Reads neither ImplCplCoef1/2 nor LDwRFlx, so a caller that has those in
// shared memory can fetch them afterwards (fewer live registers during the heavy part).
template <class Arith>
__device__ __forceinline__ void bulk_fluxes(const BulkIn &in, double sig1, BulkOut &o, BulkMid &mid, Arith &ar)
{
    using namespace sfc;
    const double LatentHeatLocal[2] = {LatentHeat, LatentHeat + LatentHeatFusion};   // :198-199
    const double z0m = RoughLength;                                                  // :194-196
    const double z0h = RoughLenHeatFactor * z0m;
    const double HumdCoef = 1.0;

    double (&Frac)[2] = mid.Frac, (&QVapSat)[2] = mid.QVapSat;
    double SfcVirTemp[2];
    Frac[0] = 1.0 - in.SIceCon;                                                      // :205-206
    Frac[1] = in.SIceCon;
    // xy_CalcFlag (:268-272): the sea-ice slot of an ice-free column.  The reference still evaluates the
    // whole slot there and multiplies by bulk coefficients that BulkCoefL82 set to 0; every output of the
    // slot is then an exact +0.  We write those zeros without evaluating (saturation exp, Ri, Louis
    // functions, three divisions): same bits for finite inputs, ~1/3 of the column's work saved where
    // there is no ice (columns of a warp share a latitude row, so the branch is warp-coherent).
    const bool ice = Frac[1] > 1e-12;
    QVapSat[1] = 0.0; SfcVirTemp[1] = 1.0;
#pragma unroll
    for (int n = 0; n < 2; n++) {                                                    // :208-213
        if (n == 1 && !ice) continue;
        QVapSat[n] = ar.div(EpsV * Es0, in.SfcPress)
                   * pexp(LatentHeatLocal[n] / GasRWet * (1.0 / 273.0 - ar.rcp(in.SfcTemp[n])));
        SfcVirTemp[n] = in.SfcTemp[n] * (1.0 + (((1.0 / EpsV) - 1.0) * QVapSat[n]));
    }
    const double VirTemp = in.SfcAirTemp * (1.0 + (((1.0 / EpsV) - 1.0) * in.QVap1));   // :215
    const double Press1 = in.SfcPress * sig1;                                        // :217
    // x**kappa = pexp(kappa*plog(x)): |kappa log x| << 1 here, so the result is within 1 ulp of the
    // correctly rounded power.  Quotients that share a divisor share its refined reciprocal (prep / div_by:
    // the bits of `a / b`, a third of the instructions).
    const double iRef = ar.prep(RefPress);
    const double Exner = ppow(ar.div_by(Press1, RefPress, iRef), GasRDry / CpDry, ar);        // :218
    const double SfcExner = ppow(ar.div_by(in.SfcPress, RefPress, iRef), GasRDry / CpDry, ar);   // :219
    const double iEx = ar.prep(Exner), iSEx = ar.prep(SfcExner);
    mid.iEx = iEx;
    mid.Exner = Exner; mid.SfcExner = SfcExner;
    const double VelAbs = ar.root(in.WindU * in.WindU + in.WindV * in.WindV);           // :221
    const double Height = in.SfcHeight + GasRDry / Grav * VirTemp * (1.0 - sig1);    // :223-224

    o.WindStressX[2] = 0.0; o.WindStressY[2] = 0.0; o.SenHFlx[2] = 0.0; o.LatHFlx[2] = 0.0;   // :226-240
    o.QVapMFlx[2] = 0.0; o.SUwRFlx[2] = 0.0; o.LUwRFlx[2] = 0.0;
    o.SfcTemp3 = 0.0; o.SfcAlbedo3 = 0.0;
    o.VelTC[2] = 0.0; o.TempTC[2] = 0.0; o.QVapTC[2] = 0.0;

    // Slot-independent pieces of the loop body (:250-266), evaluated once: the same expressions give the
    // same bits for n = 1 and n = 2.  (h+z0)/z0 also appears, negated, under the square roots of the
    // unstable branch: -(x)/z0 == -(x/z0) exactly.
    const double hzm = ar.div(Height - in.SfcHeight + z0m, z0m);
    const double hzh = (z0h == z0m) ? hzm : ar.div(Height - in.SfcHeight + z0h, z0h);
    const double lgm = plog(hzm, ar);
    const double lgh = (z0h == z0m) ? lgm : plog(hzh, ar);
    const double tmp = ar.div(FKarm, lgm);                                                 // :250-253
    const double CMn = tmp * tmp;
    const double CHn = tmp * ((z0h == z0m) ? tmp : ar.div(FKarm, lgh));                  // :255-259 (same expression, same bits)
    const double vr = fmax(VelAbs, VelMinForRi);
    const double vr2 = vr * vr;
    const double ivr2 = ar.prep(vr2);
    const double vtx = ar.div_by(VirTemp, Exner, iEx);            // slot-independent quotients of the loop body
    const double atx = ar.div_by(in.SfcAirTemp, Exner, iEx);

#pragma unroll
    for (int n = 0; n < 2; n++) {                                                    // :244
        if (n == 1 && !ice) {
            o.VelTC[n] = 0.0; o.TempTC[n] = 0.0; o.QVapTC[n] = 0.0;
            o.WindStressX[n] = 0.0; o.WindStressY[n] = 0.0; o.SenHFlx[n] = 0.0; o.QVapMFlx[n] = 0.0;
            o.LatHFlx[n] = 0.0; o.LUwRFlx[n] = 0.0; o.SUwRFlx[n] = 0.0;
            continue;
        }
        const double svx = ar.div_by(SfcVirTemp[n], SfcExner, iSEx);
        const double Ri = ar.div_by(ar.div(Grav, svx)                                // :261-266
                                    * (vtx - svx),
                                    vr2, ivr2)
                        * (Height - in.SfcHeight);
        const bool flag = (n == 0) ? true : ice;                                     // :268-272

        // ---- BulkCoefL82 (:485-568) ----
        double CM, CH, CQ;
        if (flag) {
            if (Ri > 0.0) {
                const double sq = ar.root(1.0 + 5.0 * Ri);
                CM = ar.div(CMn, 1.0 + ar.div(10.0 * Ri, sq));
                CH = ar.div(CHn, 1.0 + 15.0 * Ri * sq);
                CQ = CH;
            } else {
                CM = CMn * (1.0 - ar.div(10.0 * Ri,
                                         1.0 + 75.0 * CMn * ar.root(-hzm * Ri)));
                CH = CHn * (1.0 - ar.div(15.0 * Ri,
                                         1.0 + 75.0 * CHn * ar.root(-hzh * Ri)));
                CQ = CH;
            }
        } else {
            CM = 0.0; CH = 0.0; CQ = 0.0;
        }
        CM = fmax(fmin(CM, VelBulkCoefMax), VelBulkCoefMin);
        CH = fmax(fmin(CH, TempBulkCoefMax), TempBulkCoefMin);
        CQ = fmax(fmin(CQ, QVapBulkCoefMax), QVapBulkCoefMin);

        // ---- transfer coefficients and fluxes (:286-349) ----
        const double rt = GasRDry * SfcVirTemp[n];
        const double irt = ar.prep(rt);
        o.VelTC[n] = ar.div_by(CM * in.SfcPress, rt, irt)
                   * fmin(fmax(VelAbs, VelMinForVel), VelMaxForVel);
        o.TempTC[n] = ar.div_by(CH * in.SfcPress, rt, irt)
                    * fmin(fmax(VelAbs, VelMinForTemp), VelMaxForTemp);
        o.QVapTC[n] = ar.div_by(CQ * in.SfcPress, rt, irt)
                    * fmin(fmax(VelAbs, VelMinForQVap), VelMaxForQVap);
        if (flag) {
            o.WindStressX[n] = -o.VelTC[n] * in.WindU;
            o.WindStressY[n] = -o.VelTC[n] * in.WindV;
            o.SenHFlx[n] = -CpDry * SfcExner * o.TempTC[n]
                         * (atx - ar.div_by(in.SfcTemp[n], SfcExner, iSEx));
            o.QVapMFlx[n] = -HumdCoef * o.QVapTC[n] * (in.QVap1 - QVapSat[n]);
            o.LatHFlx[n] = LatentHeatLocal[n] * o.QVapMFlx[n];
            const double t2 = in.SfcTemp[n] * in.SfcTemp[n];
            const double t4 = t2 * t2;
            o.LUwRFlx[n] = StB * t4;
            o.SUwRFlx[n] = in.SfcAlbedo[n] * in.SDwRFlx;

            o.SfcTemp3 = o.SfcTemp3 + Frac[n] * t4;
            o.SfcAlbedo3 = o.SfcAlbedo3 + Frac[n] * in.SfcAlbedo[n];
            o.WindStressX[2] = o.WindStressX[2] + Frac[n] * o.WindStressX[n];
            o.WindStressY[2] = o.WindStressY[2] + Frac[n] * o.WindStressY[n];
            o.SenHFlx[2] = o.SenHFlx[2] + Frac[n] * o.SenHFlx[n];
            o.QVapMFlx[2] = o.QVapMFlx[2] + Frac[n] * o.QVapMFlx[n];
            o.LatHFlx[2] = o.LatHFlx[2] + Frac[n] * o.LatHFlx[n];
            o.LUwRFlx[2] = o.LUwRFlx[2] + Frac[n] * o.LUwRFlx[n];
            o.SUwRFlx[2] = o.SUwRFlx[2] + Frac[n] * o.SUwRFlx[n];
            o.VelTC[2] = o.VelTC[2] + Frac[n] * o.VelTC[n];
            o.TempTC[2] = o.TempTC[2] + Frac[n] * o.TempTC[n];
            o.QVapTC[2] = o.QVapTC[2] + Frac[n] * o.QVapTC[n];
        } else {
            o.WindStressX[n] = 0.0; o.WindStressY[n] = 0.0; o.SenHFlx[n] = 0.0; o.QVapMFlx[n] = 0.0;
            o.LatHFlx[n] = 0.0; o.LUwRFlx[n] = 0.0; o.SUwRFlx[n] = 0.0;
        }
    }
}

// Phase 2 (:353-415): implicit surface-layer update, flux correction, net heat fluxes and dF/dTs.
template <class Arith>
__device__ __forceinline__ void bulk_implicit(const BulkIn &in, const BulkMid &mid, BulkOut &o, Arith &ar)
{
    using namespace sfc;
    const double LatentHeatLocal[2] = {LatentHeat, LatentHeat + LatentHeatFusion};
    const double HumdCoef = 1.0;
    const double Exner = mid.Exner, SfcExner = mid.SfcExner;
    const double (&Frac)[2] = mid.Frac;

    // ---- implicit surface-layer update (:353-382) ----
    {
        const double DFsDT1 = ar.div_by(-CpDry * SfcExner * o.TempTC[2], Exner, mid.iEx);
        const double g0 = ar.rcp(in.Coef1[0] + o.VelTC[2]);
        const double g1 = ar.rcp(in.Coef1[1] + o.VelTC[2]);
        const double g2 = ar.rcp(in.Coef1[2] - DFsDT1);
        const double g3 = ar.rcp(in.Coef1[3] + HumdCoef * o.QVapTC[2]);
        o.Del[0] = g0 * (o.WindStressX[2] + in.Coef2[0]);
        o.Del[1] = g1 * (o.WindStressY[2] + in.Coef2[1]);
        o.Del[2] = g2 * (o.SenHFlx[2] + in.Coef2[2]);
        o.Del[3] = g3 * (o.QVapMFlx[2] + in.Coef2[3]);
        double lat3 = 0.0;
        const double cee = ar.div_by(CpDry * SfcExner, Exner, mid.iEx);
#pragma unroll
        for (int n = 0; n < 3; n++) {
            o.WindStressX[n] = o.WindStressX[n] - o.VelTC[n] * o.Del[0];
            o.WindStressY[n] = o.WindStressY[n] - o.VelTC[n] * o.Del[1];
            o.SenHFlx[n] = o.SenHFlx[n] - cee * o.TempTC[n] * o.Del[2];
            o.QVapMFlx[n] = o.QVapMFlx[n] - HumdCoef * o.QVapTC[n] * o.Del[3];
            if (n < 2) {
                o.LatHFlx[n] = LatentHeatLocal[n] * o.QVapMFlx[n];
                lat3 = lat3 + Frac[n] * o.LatHFlx[n];
            } else {
                // The reference indexes a_LatentHeatLocal(3) out of bounds here (:181,:370,:379);
                // store the area-weighted composite of the corrected slots instead.
                o.LatHFlx[n] = lat3;
            }
        }
    }

    // ---- net non-solar heat flux for ocean / sea ice (:384-400); the parts of that block that do not
    //      depend on the implicit update are in bulk_static_net ----
#pragma unroll
    for (int n = 0; n < 2; n++) {
        const bool flag = (n == 0) ? true : (Frac[n] > 1e-12);
        if (flag) o.HFlx_ns[n] = +o.LUwRFlx[n] - in.LDwRFlx + o.LatHFlx[n] + o.SenHFlx[n];
        else o.HFlx_ns[n] = 0.0;
    }
}

// The statements of the net-flux block (:384-415) that only need phase-1 values: net solar flux and
// dF/dTs.  Separate so that a caller can evaluate and store them before it fetches the phase-2 inputs.
template <class Arith>
__device__ __forceinline__ void bulk_static_net(const BulkIn &in, const BulkMid &mid, BulkOut &o, Arith &ar)
{
    using namespace sfc;
    const double LatentHeatLocal[2] = {LatentHeat, LatentHeat + LatentHeatFusion};
    const double HumdCoef = 1.0;
#pragma unroll
    for (int n = 0; n < 2; n++) {
        const bool flag = (n == 0) ? true : (mid.Frac[n] > 1e-12);
        if (flag) {
            o.HFlx_sr[n] = o.SUwRFlx[n] - in.SDwRFlx;
            const double t = in.SfcTemp[n];
            o.DHFlxDTs[n] = +4.0 * StB * (t * t * t)
                          + CpDry * o.TempTC[n]
                          + LatentHeatLocal[n] * HumdCoef * o.QVapTC[n]
                            * ar.div(LatentHeatLocal[n] * mid.QVapSat[n], GasRWet * (t * t));
        } else {
            o.HFlx_sr[n] = 0.0; o.DHFlxDTs[n] = 0.0;
        }
    }
}

// one column, fast arithmetic with the IEEE re-evaluation when a fast path was not acceptable
__device__ __forceinline__ void bulk_column(const BulkIn &in, double sig1, BulkOut &o)
{
    BulkMid mid;
    FastArith fa;
    bulk_fluxes(in, sig1, o, mid, fa);
    bulk_static_net(in, mid, o, fa);
    bulk_implicit(in, mid, o, fa);
    if (!fa.good()) {
        IeeeArith ia;
        bulk_fluxes(in, sig1, o, mid, ia);
        bulk_static_net(in, mid, o, ia);
        bulk_implicit(in, mid, o, ia);
    }
}

}  // namespace dccm
