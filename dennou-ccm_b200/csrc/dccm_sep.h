// dccm_sep.h -- separable description of a generated mapping table (host side).
//
// Every table the generators write is an outer product: the longitude part of an entry depends only on the
// destination column, the latitude part only on the destination row,
//     entry(iD, jD; m, n) = ( source cell (xi[m], yj[n]),  weight xw[m] * yw[n] ),   m in x-list(iD), n in y-list(jD).
// For grid pairs with different longitudes (ocean <-> exchange grid at BASELINE configs 3-5) the expanded table is
// O(nx*ny) entries of 12 bytes -- 0.4 GB each at config 5 -- that the kernels stream once per exchange behind two
// dependent loads (row pointer, then pairs).  The factors are O(nx + ny), live in L1/L2, and the kernels rebuild
// each entry with the generator's own multiplication (and its 1e-14 drop test), so the result has the bits of
// the expanded table (tests/test_gpu_parity.py) while the table is never built: SURVEY 8f rank 2.
#pragma once
#include <cstdint>
#include <vector>

namespace dccm {

struct SepFactors {
    bool ok = false;           // false: this grid pair / accuracy order is not handled in separable form
    int mode = 0;              // 0: Jones99 order (m outer, n inner), |w| <= 1e-14 dropped (ref jones99 :240-247)
                               // 1: bilinear order (m0,n0) (m1,n0) (m1,n1) (m0,n1), nothing dropped (ref grid_mapping_util :121-126)
    int nxs = 0, nys = 0, nxd = 0, nyd = 0;
    std::vector<int32_t> xptr, xi, yptr, yj;   // CSR lists per destination column / row, 0-based source indices
    std::vector<double> xw, yw;
    // Equal longitudes (or an axisymmetric side): the table repeats one ordered stencil (longitude shift di,
    // source row jS, weight) along each destination row -- the zonal form (kind 1), described here directly so that
    // the O(nx*ny) table need not be generated just to be recognised as a stencil afterwards.
    bool zonal = false;
    std::vector<int32_t> zptr, zdi, zjs;       // per destination row; source column = (iD + di) mod nxs
    std::vector<double> zw;
};

int jones99_factors(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                    int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                    const double *y_LatIntWtS, const double *y_LatIntWtD,
                    int accuracy_order, int lon_mode, SepFactors &f);
int bilinear_factors(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                     int nxr, const double *x_LonR, int nyr, const double *y_LatR, int lon_mode, SepFactors &f);

// Latitude band of the factors: destination rows [j0, j1) only, source rows renumbered from src_row0 (the band's
// source buffer holds src_rows rows).  Real entries must lie inside that buffer (error otherwise).
int slice_rows(SepFactors &f, int j0, int j1, int src_row0, int src_rows);

}  // namespace dccm
