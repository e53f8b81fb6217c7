"""Shared helpers of the test-suite."""
import numpy as np


def relerr(a, b, floor=0.0):
    """max over cells of |a-b| / max(|b|, floor)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if np.any(na != nb):
        return float("inf")          # NaN (= never written) in one but not the other
    a, b = np.where(na, 0.0, a), np.where(nb, 0.0, b)
    den = np.maximum(np.abs(b), floor)
    den[den == 0.0] = 1.0
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def cell_area(grid_edges_sin, im):
    """A = dlon * (sin(phi_N) - sin(phi_S)) per cell, i fastest."""
    w = np.diff(grid_edges_sin)
    return np.repeat(w, im) * (2.0 * np.pi / im)


def pair(orc, dccm, name):
    """Grid triples (ATM, OCN, SFC) of the BASELINE configs at oracle-friendly sizes,
    built by the product; the oracle gets the very same axes."""
    T = dccm.tables
    spec = {
        "T21_Pl42": (64, 32, 1, 64, False),       # shipped APEI07Couple: axisymmetric ocean
        "T42_Pl42": (128, 64, 1, 64, False),
        "T42_T42": (128, 64, 128, 64, False),     # config 1 (b)
        "T21_1deg": (64, 32, 72, 36, True),       # small mismatched-longitude case
        "T106_1deg": (320, 160, 360, 180, True),  # config 3
    }[name]
    im, jm, io, jo, reg = spec
    A = T.get_LonLatGrid(im, jm)
    O = T.regular_LonLatGrid(io, jo) if reg else T.get_LonLatGrid(io, jo)
    S = T.generate_surface_exchange_grid(A, O)
    return A, O, S


def as_orc_grid(orc, g):
    return orc.Grid(g.im, g.jm, g.x_Lon, g.y_Lat, g.x_LonWt, g.y_LatWt)
