// dccm_remap_internal.h -- device-side form of one mapping table (shared by K1 and the fused
// surface kernel).
#pragma once
#include "dccm_common.h"

struct dccm_remap {
    int n_send = 0, n_recv = 0;
    int64_t nnz = 0;
    int max_row_nnz = 0;
    int kind = 0;
    // kind 0: destination-row CSR, rows keep table order
    int32_t *d_rowptr = nullptr;   // (n_recv + 1)
    int32_t *d_col = nullptr;      // 0-based source index
    double *d_w = nullptr;
    // kind 1: zonal stencil.  Every cell (iD, jD) of destination latitude row jD applies the SAME
    // ordered list of (longitude offset di, source row jS, weight w): source cell = jS*nxs +
    // (iD + di) mod nxs.  True for every table the reference generator can produce (equal
    // longitudes or nx == 1, ref common/grid_mapping_util_jones99.f90:402-419; bilinear with equal
    // longitudes): the table shrinks from O(nx*ny) to O(ny) entries, stays in L1/L2, and the
    // gather needs no per-cell index loads at all.  Detected bit-exactly at creation.
    int nxs = 0, nxd = 0, nyd = 0;
    int64_t znnz = 0;
    int32_t *d_zptr = nullptr;     // (nyd + 1)
    int32_t *d_zdj = nullptr;      // interleaved (di, jS) pairs
    double *d_zw = nullptr;
    dccm::DevBuf send_buf, recv_buf;
};
