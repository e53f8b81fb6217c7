// dccm_remap_internal.h -- device-side form of one mapping table (shared by K1 and the fused
// surface kernel).
#pragma once
#include "dccm_common.h"

// A source buffer of a sharded run as the kernels see it: its first b0 cells are the lower
// neighbour's boundary rows and the cells from b1 on the upper neighbour's -- read IN PLACE from
// the neighbours' own send buffers over NVLink peer mappings (no halo copy, no collective); all three
// bases are pre-offset so that element (cell c, layer l) is base[c + l*ld].  Unsharded: b0 = 0,
// b1 = INT64_MAX and everything resolves to `own`.
struct SrcSeg {
    const double *lo, *own, *hi;
    int64_t b0, b1;
#ifdef __CUDACC__
    __device__ __forceinline__ const double *at(int64_t c) const { return (c < b0 ? lo : (c >= b1 ? hi : own)) + c; }
#endif
};

inline SrcSeg seg_of(const double *p) { return SrcSeg{p, p, p, 0, INT64_MAX}; }
inline SrcSeg seg_of(const dccm_src_seg *s) { return SrcSeg{s->lo, s->own, s->hi, s->b0, s->b1}; }

// kind-2 table as the kernels see it: fixed-width lists, wx entries per destination column and wy per destination
// row (shorter lists are padded with weight 0, which the mode-0 drop test |w| > 1e-14 discards like the generator
// discards any other negligible entry; mode 1 lists are exactly 2 wide), so a cell finds its entries at
// iD*wx / jD*wy without a pointer load.
struct SepTab {
    const int32_t *xi, *yj;
    const double *xw, *yw;
    int mode, nxs, nxd, wx, wy;
};

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
    // what a CTA that stages source tiles in shared memory must provision (TMA path of the fused
    // surface kernel): longest stencil, most distinct source rows in one stencil, and the range of
    // the longitude offsets taken as signed shifts (di > nxs/2 is a westward neighbour)
    int z_max_len = 0, z_max_rows = 0, z_dmin = 0, z_dmax = 0;
    // kind 2: separable form of a generated table (dccm_sep.h): per destination column a list of (source column,
    // longitude factor), per destination row a list of (source row, latitude factor); the kernels rebuild each
    // entry as the generator did (weight = xw * yw, mode 0 drops |w| <= 1e-14).  nxs / nxd / nyd as for kind 1.
    int sep_mode = 0, sep_wx = 0, sep_wy = 0;
    int32_t *d_xi = nullptr, *d_yj = nullptr;
    double *d_xw = nullptr, *d_yw = nullptr;
    // fused surface kernel: cells to re-evaluate with plain IEEE operators (csrc/dccm_exchange.cu); the list of
    // the A->S bilinear handle is the one used.  Allocated at creation (never inside a stream capture).
    static constexpr int kRedoCap = 16384;
    int *d_redo = nullptr;         // [count, done, (member, cell) x kRedoCap]
    // options / bookkeeping of the fused surface kernel, kept on the A->S bilinear handle of the call (per handle, so
    // two exchanges configured differently do not see each other): form requested (1 staged, 0 direct), CTAs per SM
    // the kernel is built for (4, 5 or 6) and the form the last call on this handle took (-1: none yet)
    int sfc_staged = 1, sfc_minb = 6, sfc_last_form = -1;      // 6 CTAs/SM (80 registers): the kernel is latency bound, warps beat spills (profiles/r02d)
    dccm::DevBuf send_buf, recv_buf;
};
