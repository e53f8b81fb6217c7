"""Seed-free synthetic coupled fields for tests and bench.py.

Every value is a pure function of the GLOBAL cell index (hash noise), so a field is identical
no matter how the grid is sharded over ranks, and can be produced either with numpy (tests,
CPU baseline) or directly on the device with torch (bench at full size).  `xp` is the array
module: numpy or torch; `dev` the torch device (ignored for numpy).

All arrays use the reversed-axis convention of the host mirrors: (slot/level, cell).
"""
import math

import numpy as np

# DCPAM-side constants handed to the implicit coupling solve (ref atm/dcpam_sfc_implicit_coupling_mod.f90:85-108
# takes them from DCPAM's `constants`; values as in exp/APEI07Couple/common/dcpam/*.conf defaults)
GRAV = 9.8
CPDRY = 1004.6
GASRDRY = 287.04
DELTIME = 1200.0
SIG1 = 0.995          # a_Sig1Info(1), SURVEY 8d


def _is_torch(xp):
    return xp.__name__ == "torch"


def _arange(xp, lo, hi, dev):
    if _is_torch(xp):
        return xp.arange(lo, hi, dtype=xp.float64, device=dev)
    return xp.arange(lo, hi, dtype=xp.float64)


def _asx(xp, a, dev):
    if _is_torch(xp):
        return xp.as_tensor(np.asarray(a, dtype=np.float64), dtype=xp.float64, device=dev)
    return np.asarray(a, dtype=np.float64)


def unit(xp, idx, salt):
    """uniform (0,1) hash noise of the global index."""
    v = xp.sin(idx * 12.9898 + salt * 78.233 + 0.1) * 43758.5453
    return v - xp.floor(v)


def normal(xp, idx, salt):
    """approx N(0,1): sum of 4 uniforms, rescaled."""
    s = unit(xp, idx, salt) + unit(xp, idx, salt + 0.37) + unit(xp, idx, salt + 0.71) + unit(xp, idx, salt + 1.13)
    return (s - 2.0) * math.sqrt(3.0)


def cell_coords(xp, grid, j0=0, j1=None, dev=None):
    """global linear index, lon, lat (1-D, i fastest) of the latitude band [j0, j1)."""
    j1 = grid.jm if j1 is None else j1
    lon = _asx(xp, grid.x_Lon, dev)
    lat = _asx(xp, grid.y_Lat[j0:j1], dev)
    n = (j1 - j0) * grid.im
    idx = _arange(xp, j0 * grid.im, j0 * grid.im + n, dev)
    if _is_torch(xp):
        lo = lon.repeat(j1 - j0)
        la = lat.repeat_interleave(grid.im)
    else:
        lo = np.tile(lon, j1 - j0)
        la = np.repeat(lat, grid.im)
    return idx, lo, la


def generic_fields(xp, grid, D, j0=0, j1=None, dev=None, salt=0.0):
    """field d = sin(d*lon)cos(lat) + 0.1 N(0,1)   (SURVEY 8d), shape (D, n)."""
    idx, lon, lat = cell_coords(xp, grid, j0, j1, dev)
    out = [xp.sin((d + 1) * lon) * xp.cos(lat) + 0.1 * normal(xp, idx, salt + 3.0 * d) for d in range(D)]
    return xp.stack(out)


def qsat(xp, T, P):
    return 611.0 / P * xp.exp(2425300.0 / (8.3144621 / 0.018) * (1.0 / 273.0 - 1.0 / T))


def atm_surface_fields(xp, grid, j0=0, j1=None, dev=None, member=0):
    """The a2s fields other than the coupling coefficients, on the ATM grid: dict of (n,) arrays."""
    idx, lon, lat = cell_coords(xp, grid, j0, j1, dev)
    s = 100.0 * member
    T1 = 288.0 - 40.0 * xp.sin(lat) ** 2 + 2.0 * normal(xp, idx, s + 1.0)
    Ps = 1.0e5 + 500.0 * normal(xp, idx, s + 2.0)
    f = {
        "WindU": 8.0 * normal(xp, idx, s + 3.0),
        "WindV": 8.0 * normal(xp, idx, s + 4.0),
        "SfcAirTemp": T1,
        "SfcPress": Ps,
        "QVap1": 0.8 * qsat(xp, T1, Ps),
        "LDwRFlx": 300.0 + 20.0 * normal(xp, idx, s + 5.0),
        "SDwRFlx": 340.0 * (1.0 + 0.01 * member) * xp.cos(lat) * (0.5 + 0.5 * unit(xp, idx, s + 6.0)),
        "RainFall": -3e-5 * xp.log(1.0 - 0.999 * unit(xp, idx, s + 7.0)),
        "SnowFall": -3e-5 * xp.log(1.0 - 0.999 * unit(xp, idx, s + 8.0)),
    }
    return f


def ocn_surface_fields(xp, grid, j0=0, j1=None, dev=None, member=0):
    """o2s / i2s fields on the OCN grid: SfcTemp(2), SfcAlbedo(2), SIceCon."""
    idx, lon, lat = cell_coords(xp, grid, j0, j1, dev)
    s = 100.0 * member + 50.0
    T1 = 288.0 - 40.0 * xp.sin(lat) ** 2
    Ts_o = T1 + 1.5 * normal(xp, idx, s + 1.0)
    Ts_o = xp.clip(Ts_o, 271.35, None) if not _is_torch(xp) else xp.clamp(Ts_o, min=271.35)
    Ts_i = Ts_o - 5.0
    Ts_i = xp.clip(Ts_i, None, 273.15) if not _is_torch(xp) else xp.clamp(Ts_i, max=273.15)
    ice = (xp.abs(lat) * (180.0 / math.pi) - 60.0) / 20.0
    ice = xp.clip(ice, 0.0, 1.0) if not _is_torch(xp) else xp.clamp(ice, 0.0, 1.0)
    return {
        "SfcTempO": Ts_o, "SfcTempI": Ts_i,
        "SfcAlbedoO": 0.1 + 0.0 * lat, "SfcAlbedoI": 0.6 + 0.0 * lat,
        "SIceCon": ice,
    }


def sigma_half(K):
    """sigma at half levels k = 0..K: 1 at the surface, 0.99 at k = 1 (first full level ~0.995,
    as DCPAM's lowest level), quadratic to 0 at the top."""
    s = np.zeros(K + 1)
    s[0] = 1.0
    for k in range(1, K + 1):
        s[k] = 0.99 * (1.0 - (k - 1.0) / (K - 1.0)) ** 2 if K > 1 else 0.0
    s[K] = 0.0
    return s


def column_inputs(xp, grid, K, ncmax=1, j0=0, j1=None, dev=None, member=0):
    """Inputs of SfcImplicitCoupling_VDiffForward for the band: dict keyed like IN_ORDER.
    Half-level arrays (K+1, n), full-level arrays (K, n), tracer arrays (ncmax, levels, n)."""
    idx, lon, lat = cell_coords(xp, grid, j0, j1, dev)
    s = 100.0 * member + 20.0
    kappa = GASRDRY / CPDRY
    sh = sigma_half(K)
    sf = 0.5 * (sh[:-1] + sh[1:])
    sf[-1] = 0.5 * sh[-2]
    Ps = 1.0e5 + 500.0 * normal(xp, idx, 2.0 + 100.0 * member)
    T1 = 288.0 - 40.0 * xp.sin(lat) ** 2 + 2.0 * normal(xp, idx, 1.0 + 100.0 * member)
    U1 = 8.0 * normal(xp, idx, 3.0 + 100.0 * member)
    V1 = 8.0 * normal(xp, idx, 4.0 + 100.0 * member)
    Hs = GASRDRY * 260.0 / GRAV
    zf = [-Hs * math.log(sf[k]) for k in range(K)]                 # full-level heights (scalars)
    zh = [-Hs * math.log(max(sh[k], 0.5 * sf[-1])) for k in range(K + 1)]

    def stack(lst):
        return xp.stack(lst)

    wob = [1.0 + 0.002 * normal(xp, idx, s + 0.01 * k) for k in range(K + 1)]
    Press = stack([Ps * sh[k] for k in range(K + 1)])
    Height = stack([zf[k] * wob[k] for k in range(K)])
    Tfull = [xp.clip(T1 - 6.5e-3 * zf[k], 200.0, None) if not _is_torch(xp) else xp.clamp(T1 - 6.5e-3 * zf[k], min=200.0)
             for k in range(K)]
    Thalf = [Tfull[0]] + [0.5 * (Tfull[k - 1] + Tfull[k]) for k in range(1, K)] + [Tfull[K - 1]]
    VirTemp = stack([Thalf[k] * (1.0 + 0.005 * unit(xp, idx, s + 5.0 + k)) for k in range(K + 1)])
    zExner = stack([(Ps * sf[k] / 1.0e5) ** kappa for k in range(K)])
    rExner = stack([(Ps * sh[k] / 1.0e5) ** kappa for k in range(K + 1)])
    diff = [10.0 * math.exp(-zh[k] / 1000.0) + 0.1 for k in range(K + 1)]
    VelDiff = stack([diff[k] * (1.0 + 0.2 * unit(xp, idx, s + 30.0 + k)) for k in range(K + 1)])
    TempDiff = stack([1.2 * diff[k] * (1.0 + 0.2 * unit(xp, idx, s + 60.0 + k)) for k in range(K + 1)])
    QMixDiff = stack([1.1 * diff[k] * (1.0 + 0.2 * unit(xp, idx, s + 90.0 + k)) for k in range(K + 1)])

    # fluxes: -rho*D*dX/dz between neighbouring full levels, surface flux from a bulk formula, 0 at the top
    def flux(prof, surf, dcoef):
        out = [surf]
        for k in range(1, K):
            rho = Press[k] / (GASRDRY * VirTemp[k])
            out.append(-rho * dcoef[k] * (prof[k] - prof[k - 1]) / (Height[k] - Height[k - 1]))
        out.append(0.0 * surf)
        return stack(out)

    Uprof = [U1 * (1.0 + 0.4 * math.log1p(zf[k] / 50.0)) + 0.5 * normal(xp, idx, s + 120.0 + k) for k in range(K)]
    Vprof = [V1 * (1.0 + 0.4 * math.log1p(zf[k] / 50.0)) + 0.5 * normal(xp, idx, s + 150.0 + k) for k in range(K)]
    Th = [Tfull[k] / zExner[k] + 0.3 * normal(xp, idx, s + 180.0 + k) for k in range(K)]
    spd = xp.sqrt(U1 * U1 + V1 * V1) + 0.1
    rho_s = Ps / (GASRDRY * T1)
    MomFluxX = flux(Uprof, -1.3e-3 * rho_s * spd * U1, VelDiff)
    MomFluxY = flux(Vprof, -1.3e-3 * rho_s * spd * V1, VelDiff)
    HeatFlux = CPDRY * flux(Th, 15.0 / CPDRY * (1.0 + normal(xp, idx, s + 7.0)), TempDiff)
    q1 = 0.8 * qsat(xp, T1, Ps)
    QF = []
    for n in range(ncmax):
        Qprof = [q1 * math.exp(-zf[k] / (2500.0 + 500.0 * n)) * (1.0 + 0.05 * normal(xp, idx, s + 210.0 + k + 40 * n))
                 for k in range(K)]
        QF.append(flux(Qprof, 3e-5 * (1.0 + 0.5 * normal(xp, idx, s + 8.0 + n)), QMixDiff))
    return {
        "MomFluxX": MomFluxX, "MomFluxY": MomFluxY, "HeatFlux": HeatFlux, "QMixFlux": stack(QF),
        "Press": Press, "zExner": zExner, "rExner": rExner, "VirTemp": VirTemp, "Height": Height,
        "VelDiffCoef": VelDiff, "TempDiffCoef": TempDiff, "QMixDiffCoef": QMixDiff,
    }
