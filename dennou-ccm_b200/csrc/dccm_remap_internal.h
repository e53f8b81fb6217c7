// dccm_remap_internal.h -- device-side form of one mapping table (shared by K1 and the fused
// surface kernel).
#pragma once
#include "dccm_common.h"

struct dccm_remap {
    int n_send = 0, n_recv = 0;
    int64_t nnz = 0;
    int max_row_nnz = 0;
    int kind = 0;
    int32_t *d_rowptr = nullptr;   // (n_recv + 1) destination-row CSR, rows keep table order
    int32_t *d_col = nullptr;      // 0-based source index
    double *d_w = nullptr;
    dccm::DevBuf send_buf, recv_buf;
};
