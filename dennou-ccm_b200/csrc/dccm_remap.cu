// dccm_remap.cu -- K1: remap apply (batched multi-field gather-SpMV).
//
// Replaces interpolate_data_latlon (ref common/interpolation_data_latlon_mod.f90:274-306):
//     recv(:,:) = 0 ; do d ; do i : recv(r_i,d) = recv(r_i,d) + send(s_i,d)*coefS(i)
// The COO operation list is turned ONCE (handle creation) into destination-row CSR with a
// stable counting sort, so every destination row keeps its operations in table order and is
// accumulated by a single thread in exactly the reference's summation order -- with separate
// IEEE multiply and add (no FMA contraction) the result is bit-identical to the reference
// loop, while the scatter / read-modify-write and the per-field re-read of the index and
// coefficient arrays are gone: indices+weights are read once for all fields of a call.
//
// Layout: send(sn1, nfield), recv(rn1, nfield) column-major (point fastest, field slowest) is
// fixed by the Jcup boundary.  Rows are latitude-major and consecutive rows are consecutive
// longitudes, so with one thread per row a warp's gathers hit consecutive source points.
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

#include "dccm_common.h"

using namespace dccm;

#include "dccm_remap_internal.h"
#include "dccm_sep.h"

SepTab sep_of(const dccm_remap *h);

namespace {

constexpr int kThreads = 256;

template <bool SEG>
__device__ __forceinline__ const double *cell(const SrcSeg &s, int64_t c) { return SEG ? s.at(c) : s.own + c; }

// One thread per destination row; FB fields register-blocked (the launcher picks the smallest FB that
// takes all fields of the call in one pass, so a row's (col, w) pairs are read once).  The pairs are
// fetched CH at a time before the dependent source loads are issued: a thread has up to CH*FB gathers
// in flight (rows hold 1-6 entries).  Accumulation stays in table order, multiply and add separate.
template <int FB, bool SEG>
__global__ void __launch_bounds__(kThreads)
remap_csr_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                 const double *__restrict__ w, const SrcSeg send, int64_t sn1,
                 double *__restrict__ recv, int64_t rn1, int n_recv, int nfield, int fields_per_y)
{
    constexpr int CH = FB <= 5 ? 4 : 2;
    const int r = blockIdx.x * kThreads + threadIdx.x;
    if (r >= n_recv) return;
    const int k0 = __ldg(&rowptr[r]), k1 = __ldg(&rowptr[r + 1]);
    const int d_begin = blockIdx.y * fields_per_y;
    const int d_end = min(nfield, d_begin + fields_per_y);
    for (int d0 = d_begin; d0 < d_end; d0 += FB) {
        double acc[FB];
#pragma unroll
        for (int d = 0; d < FB; d++) acc[d] = 0.0;
        const int64_t o0 = (int64_t)d0 * sn1;
        const int nf = min(FB, d_end - d0);
        for (int kb = k0; kb < k1; kb += CH) {
            int c[CH];
            double ww[CH];
#pragma unroll
            for (int j = 0; j < CH; j++) {
                const bool on = kb + j < k1;
                c[j] = on ? __ldg(&col[kb + j]) : 0;
                ww[j] = on ? __ldg(&w[kb + j]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (kb + j < k1) {
                    const double *sp = cell<SEG>(send, c[j]) + o0;
                    if (nf == FB) {
#pragma unroll
                        for (int d = 0; d < FB; d++)
                            acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), ww[j]));
                    } else {
#pragma unroll
                        for (int d = 0; d < FB; d++)
                            if (d < nf) acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), ww[j]));
                    }
                }
            }
        }
#pragma unroll
        for (int d = 0; d < FB; d++)
            if (d < nf) recv[r + (int64_t)(d0 + d) * rn1] = acc[d];
    }
}

// kind 1: zonal stencil, one thread per destination cell, no per-cell table reads.
template <int FB, bool SEG>
__global__ void __launch_bounds__(kThreads)
remap_zonal_kernel(const int32_t *__restrict__ zptr, const int32_t *__restrict__ zdj,
                   const double *__restrict__ zw, int nxs, int nxd,
                   const SrcSeg send, int64_t sn1, double *__restrict__ recv, int64_t rn1,
                   int n_recv, int nfield, int fields_per_y)
{
    const int r = blockIdx.x * kThreads + threadIdx.x;
    if (r >= n_recv) return;
    const int jD = r / nxd, iD = r - jD * nxd;
    const int e0 = __ldg(&zptr[jD]), e1 = __ldg(&zptr[jD + 1]);
    const int d_begin = blockIdx.y * fields_per_y;
    const int d_end = min(nfield, d_begin + fields_per_y);
    for (int d0 = d_begin; d0 < d_end; d0 += FB) {
        double acc[FB];
#pragma unroll
        for (int d = 0; d < FB; d++) acc[d] = 0.0;
        const int64_t o0 = (int64_t)d0 * sn1;
        const int nf = min(FB, d_end - d0);
        for (int e = e0; e < e1; e++) {
            int i = iD + __ldg(&zdj[2 * e]);
            if (i >= nxs) i -= nxs;
            if (nxs == 1) i = 0;                     // axisymmetric source: every longitude reads column 1
            const int64_t c = (int64_t)__ldg(&zdj[2 * e + 1]) * nxs + i;
            const double ww = __ldg(&zw[e]);
            const double *sp = cell<SEG>(send, c) + o0;
            if (nf == FB) {
#pragma unroll
                for (int d = 0; d < FB; d++)
                    acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), ww));
            } else {
#pragma unroll
                for (int d = 0; d < FB; d++)
                    if (d < nf) acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), ww));
            }
        }
#pragma unroll
        for (int d = 0; d < FB; d++)
            if (d < nf) recv[r + (int64_t)(d0 + d) * rn1] = acc[d];
    }
}

// kind 2: separable generated table, one thread per destination cell.  The x-list of the cell's column and the
// y-list of its row are a few L1-resident entries at fixed offsets; each (m, n) pair is rebuilt exactly as the
// generator emitted it (order, product, drop test), three latitude entries at a time so their source loads are in
// flight together.
template <int FB, bool SEG>
__global__ void __launch_bounds__(kThreads)
remap_sep_kernel(const SepTab t, const SrcSeg send, int64_t sn1, double *__restrict__ recv, int64_t rn1,
                 int n_recv, int nfield, int fields_per_y)
{
    const int r = blockIdx.x * kThreads + threadIdx.x;
    if (r >= n_recv) return;
    const int jD = r / t.nxd, iD = r - jD * t.nxd;
    const int x0 = iD * t.wx, y0 = jD * t.wy, y1 = y0 + t.wy;
    const int d_begin = blockIdx.y * fields_per_y;
    const int d_end = min(nfield, d_begin + fields_per_y);
    for (int d0 = d_begin; d0 < d_end; d0 += FB) {
        double acc[FB];
#pragma unroll
        for (int d = 0; d < FB; d++) acc[d] = 0.0;
        const int64_t o0 = (int64_t)d0 * sn1;
        const int nf = min(FB, d_end - d0);
        auto add = [&](int c, double w) {
            const double *sp = cell<SEG>(send, c) + o0;
            if (nf == FB) {
#pragma unroll
                for (int d = 0; d < FB; d++) acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), w));
            } else {
#pragma unroll
                for (int d = 0; d < FB; d++)
                    if (d < nf) acc[d] = __dadd_rn(acc[d], __dmul_rn(__ldg(sp + (int64_t)d * sn1), w));
            }
        };
        if (t.mode == 1) {                         // bilinear: (m0,n0) (m1,n0) (m1,n1) (m0,n1), nothing dropped
            const int i0 = __ldg(&t.xi[x0]), i1 = __ldg(&t.xi[x0 + 1]);
            const double a0 = __ldg(&t.xw[x0]), a1 = __ldg(&t.xw[x0 + 1]);
            const int j0 = __ldg(&t.yj[y0]) * t.nxs, j1 = __ldg(&t.yj[y0 + 1]) * t.nxs;
            const double b0 = __ldg(&t.yw[y0]), b1 = __ldg(&t.yw[y0 + 1]);
            add(j0 + i0, __dmul_rn(a0, b0)); add(j0 + i1, __dmul_rn(a1, b0));
            add(j1 + i1, __dmul_rn(a1, b1)); add(j1 + i0, __dmul_rn(a0, b1));
        } else {
            for (int m = x0; m < x0 + t.wx; m++) {
                const int i = __ldg(&t.xi[m]);
                const double a = __ldg(&t.xw[m]);
                for (int nb = y0; nb < y1; nb += 3) {
                    int c[3];
                    double w[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const bool on = nb + j < y1;
                        c[j] = on ? __ldg(&t.yj[nb + j]) * t.nxs + i : 0;
                        w[j] = on ? __dmul_rn(a, __ldg(&t.yw[nb + j])) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 3; j++)
                        if (fabs(w[j]) > 1e-14) add(c[j], w[j]);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < FB; d++)
            if (d < nf) recv[r + (int64_t)(d0 + d) * rn1] = acc[d];
    }
}

// Is the (row-sorted) table a zonal stencil on an (nxs x .) -> (nxd x nyd) grid pair?  Exact test:
// every row of a destination latitude must repeat row iD = 0 shifted in longitude with equal weights.
bool detect_zonal(const std::vector<int32_t> &rowptr, const std::vector<int32_t> &col, const std::vector<double> &w,
                  int nxs, int nxd, int nyd, std::vector<int32_t> &zptr, std::vector<int32_t> &zdi,
                  std::vector<int32_t> &zjs, std::vector<double> &zw)
{
    if (!(nxs == nxd || nxs == 1) || nxd < 2) return false;
    zptr.assign(nyd + 1, 0);
    zdi.clear(); zjs.clear(); zw.clear();
    for (int jD = 0; jD < nyd; jD++) {
        const int r0 = jD * nxd;
        const int k0 = rowptr[r0], n = rowptr[r0 + 1] - k0;
        if (n > 64) return false;
        for (int k = 0; k < n; k++) {
            zdi.push_back(col[k0 + k] % nxs);
            zjs.push_back(col[k0 + k] / nxs);
            zw.push_back(w[k0 + k]);
        }
        zptr[jD + 1] = (int32_t)zdi.size();
        const int32_t *di = zdi.data() + zptr[jD], *js = zjs.data() + zptr[jD];
        const double *ww = zw.data() + zptr[jD];
        for (int iD = 1; iD < nxd; iD++) {
            const int k1 = rowptr[r0 + iD];
            if (rowptr[r0 + iD + 1] - k1 != n) return false;
            for (int k = 0; k < n; k++) {
                int i = iD + di[k];
                if (i >= nxs) i -= nxs;
                if (nxs == 1) i = 0;
                if (col[k1 + k] != js[k] * nxs + i) return false;
                // value equality: +0.0 and -0.0 weights are interchangeable (acc starts at +0.0 and
                // x + (+-0) == x), anything else must match exactly; NaN never matches
                if (!(w[k1 + k] == ww[k])) return false;
            }
        }
    }
    return true;
}

std::mutex g_reg_mutex;
std::map<std::tuple<int, int, int>, dccm_remap *> g_registry;
std::map<std::string, int> g_model_ids;          // component name -> Jcup component number (guarded by g_reg_mutex)

// Fortran CHARACTER(*) argument -> std::string: `len` characters, trailing blanks (and a stray NUL) dropped.
std::string fortran_name(const char *s, int64_t len)
{
    if (!s || len <= 0) return std::string();
    size_t n = (size_t)len;
    if (const void *z = memchr(s, 0, n)) n = (size_t)((const char *)z - s);
    while (n > 0 && s[n - 1] == ' ') n--;
    return std::string(s, n);
}

}  // namespace

namespace {
// COO (1-based, operation order) -> destination-row CSR by a stable counting sort: per-row order == table order
int build_csr(int64_t nops, const int32_t *send_index, const int32_t *recv_index, const double *coef,
              int n_send, int n_recv, std::vector<int32_t> &rowptr, std::vector<int32_t> &col,
              std::vector<double> &w, int &maxnnz)
{
    rowptr.assign((size_t)n_recv + 1, 0);
    for (int64_t i = 0; i < nops; i++) {
        int32_t r = recv_index[i], s = send_index[i];
        if (r < 1 || r > n_recv || s < 1 || s > n_send)
            return fail(DCCM_ERR_ARG, "dccm_remap_create: op %lld has index out of range (send %d/%d, recv %d/%d)",
                        (long long)i, s, n_send, r, n_recv);
        rowptr[r]++;
    }
    maxnnz = 0;
    for (int r = 0; r < n_recv; r++) {
        maxnnz = std::max(maxnnz, rowptr[r + 1]);
        rowptr[r + 1] += rowptr[r];
    }
    col.resize((size_t)nops);
    w.resize((size_t)nops);
    std::vector<int32_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (int64_t i = 0; i < nops; i++) {
        int32_t p = fill[recv_index[i] - 1]++;
        col[p] = send_index[i] - 1;
        w[p] = coef[i];
    }
    return DCCM_OK;
}
}  // namespace

namespace {
// kind-1 storage of a handle: upload the per-row stencils and note what a CTA that stages source tiles in shared
// memory must provision (longest stencil, most distinct source rows in one stencil, range of the signed shifts)
int finish_zonal(dccm_remap *h, int nxs, int nxd, int nyd, const std::vector<int32_t> &zptr,
                 const std::vector<int32_t> &zdi, const std::vector<int32_t> &zjs, const std::vector<double> &zw)
{
    h->kind = 1; h->nxs = nxs; h->nxd = nxd; h->nyd = nyd; h->znnz = (int64_t)zw.size();
    h->z_dmin = INT32_MAX; h->z_dmax = INT32_MIN;
    for (int jD = 0; jD < nyd; jD++) {
        const int e0 = zptr[jD], e1 = zptr[jD + 1];
        h->z_max_len = std::max(h->z_max_len, e1 - e0);
        int rows = 0;
        for (int e = e0; e < e1; e++) {
            bool first = true;
            for (int f = e0; f < e; f++) first = first && zjs[f] != zjs[e];
            rows += first;
            const int sd = zdi[e] > nxs / 2 ? zdi[e] - nxs : zdi[e];
            h->z_dmin = std::min(h->z_dmin, sd); h->z_dmax = std::max(h->z_dmax, sd);
        }
        h->z_max_rows = std::max(h->z_max_rows, rows);
    }
    if (zw.empty()) h->z_dmin = h->z_dmax = 0;
    std::vector<int32_t> zdj(2 * zdi.size());
    for (size_t k = 0; k < zdi.size(); k++) { zdj[2 * k] = zdi[k]; zdj[2 * k + 1] = zjs[k]; }
    cudaError_t e = cudaMalloc(&h->d_zptr, sizeof(int32_t) * zptr.size());
    if (e == cudaSuccess) e = cudaMalloc(&h->d_zdj, sizeof(int32_t) * std::max<size_t>(2, zdj.size()));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_zw, sizeof(double) * std::max<size_t>(1, zw.size()));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_zptr, zptr.data(), sizeof(int32_t) * zptr.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_zdj, zdj.data(), sizeof(int32_t) * zdj.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_zw, zw.data(), sizeof(double) * zw.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(DCCM_ERR_CUDA, "dccm_remap_create: %s", cudaGetErrorString(e));
    return DCCM_OK;
}
}  // namespace

// Host-only: which storage form would dccm_remap_create_lonlat pick (no GPU needed).
extern "C" int dccm_remap_classify(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                                   const double *coef, int n_send, int n_recv, int gnxs, int gnxr,
                                   int *kind, int64_t *stored_entries)
{
    std::vector<int32_t> rowptr, col, zptr, zdi, zjs;
    std::vector<double> w, zw;
    int maxnnz = 0;
    if (nops < 0 || n_send < 1 || n_recv < 1) return fail(DCCM_ERR_ARG, "dccm_remap_classify: bad sizes");
    int rc = build_csr(nops, send_index, recv_index, coef, n_send, n_recv, rowptr, col, w, maxnnz);
    if (rc) return rc;
    *kind = 0; *stored_entries = nops;
    if (gnxs > 0 && gnxr > 0 && n_send % gnxs == 0 && n_recv % gnxr == 0 && nops > 0 &&
        detect_zonal(rowptr, col, w, gnxs, gnxr, n_recv / gnxr, zptr, zdi, zjs, zw)) {
        *kind = 1; *stored_entries = (int64_t)zw.size();
    }
    return DCCM_OK;
}

extern "C" int dccm_remap_create(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                                 const double *coef, int n_send, int n_recv, dccm_remap **out)
{
    return dccm_remap_create_lonlat(nops, send_index, recv_index, coef, n_send, n_recv, 0, 0, out);
}

extern "C" int dccm_remap_create_lonlat(int64_t nops, const int32_t *send_index, const int32_t *recv_index,
                                        const double *coef, int n_send, int n_recv, int gnxs, int gnxr,
                                        dccm_remap **out)
{
    *out = nullptr;
    if (nops < 0 || n_send < 1 || n_recv < 1) return fail(DCCM_ERR_ARG, "dccm_remap_create: bad sizes");
    if (nops >= INT32_MAX) return fail(DCCM_ERR_ARG, "dccm_remap_create: nops exceeds int32");
    int rc = ensure_device();
    if (rc) return rc;
    std::vector<int32_t> rowptr, col;
    std::vector<double> w;
    int maxnnz = 0;
    rc = build_csr(nops, send_index, recv_index, coef, n_send, n_recv, rowptr, col, w, maxnnz);
    if (rc) return rc;
    dccm_remap *h = new dccm_remap();
    h->n_send = n_send; h->n_recv = n_recv; h->nnz = nops; h->max_row_nnz = maxnnz;
    cudaError_t e = cudaMalloc(&h->d_redo, sizeof(int) * (2 + 2 * (size_t)dccm_remap::kRedoCap));
    if (e == cudaSuccess) e = cudaMemset(h->d_redo, 0, sizeof(int) * 2);
    if (e != cudaSuccess) {
        dccm_remap_destroy(h);
        return fail(DCCM_ERR_CUDA, "dccm_remap_create: %s", cudaGetErrorString(e));
    }
    if (gnxs > 0 && gnxr > 0 && n_send % gnxs == 0 && n_recv % gnxr == 0 && nops > 0) {
        std::vector<int32_t> zptr, zdi, zjs;
        std::vector<double> zw;
        if (detect_zonal(rowptr, col, w, gnxs, gnxr, n_recv / gnxr, zptr, zdi, zjs, zw)) {
            rc = finish_zonal(h, gnxs, gnxr, n_recv / gnxr, zptr, zdi, zjs, zw);
            if (rc) { dccm_remap_destroy(h); return rc; }
            *out = h;
            return DCCM_OK;
        }
    }
    e = cudaMalloc(&h->d_rowptr, sizeof(int32_t) * ((size_t)n_recv + 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_col, sizeof(int32_t) * std::max<size_t>(1, (size_t)nops));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_w, sizeof(double) * std::max<size_t>(1, (size_t)nops));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_rowptr, rowptr.data(), sizeof(int32_t) * rowptr.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nops) e = cudaMemcpy(h->d_col, col.data(), sizeof(int32_t) * col.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nops) e = cudaMemcpy(h->d_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        dccm_remap_destroy(h);
        return fail(DCCM_ERR_CUDA, "dccm_remap_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return DCCM_OK;
}

SepTab sep_of(const dccm_remap *h)
{
    return SepTab{h->d_xi, h->d_yj, h->d_xw, h->d_yw, h->sep_mode, h->nxs, h->nxd, h->sep_wx, h->sep_wy};
}

namespace {
template <class T> cudaError_t upload(T *&d, const std::vector<T> &v)
{
    cudaError_t e = cudaMalloc(&d, sizeof(T) * std::max<size_t>(2, v.size()));
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return e;
}

// kind-2 handle from the factors; nnz = entries the expanded table would hold (mode 0 applies the drop test)
int create_separable(const SepFactors &f, dccm_remap **out)
{
    int rc = ensure_device();
    if (rc) return rc;
    dccm_remap *h = new dccm_remap();
    h->kind = 2; h->sep_mode = f.mode; h->nxs = f.nxs; h->nxd = f.nxd; h->nyd = f.nyd;
    h->n_send = f.nxs * f.nys; h->n_recv = f.nxd * f.nyd;
    int64_t nnz = 0;
    int maxrow = 0;
    if (f.mode == 1) { nnz = 4 * (int64_t)h->n_recv; maxrow = 4; }
    else {
        // count per (x-entry, y-entry) pair once: the kept pairs of column iD and row jD multiply out
        for (int jD = 0; jD < f.nyd; jD++) {
            for (int iD = 0; iD < f.nxd; iD++) {
                int k = 0;
                for (int m = f.xptr[iD]; m < f.xptr[iD + 1]; m++)
                    for (int n = f.yptr[jD]; n < f.yptr[jD + 1]; n++)
                        k += std::fabs(f.xw[m] * f.yw[n]) > 1e-14;
                nnz += k; maxrow = std::max(maxrow, k);
            }
        }
    }
    h->nnz = nnz; h->max_row_nnz = maxrow;
    // fixed-width lists: pad with (index 0, weight 0.0)
    auto pad = [](const std::vector<int32_t> &ptr, const std::vector<int32_t> &idx, const std::vector<double> &w, int n,
                  int &width, std::vector<int32_t> &pidx, std::vector<double> &pw) {
        width = 1;
        for (int k = 0; k < n; k++) width = std::max(width, ptr[k + 1] - ptr[k]);
        pidx.assign((size_t)n * width, 0);
        pw.assign((size_t)n * width, 0.0);
        for (int k = 0; k < n; k++)
            for (int e = ptr[k]; e < ptr[k + 1]; e++) {
                pidx[(size_t)k * width + (e - ptr[k])] = idx[e];
                pw[(size_t)k * width + (e - ptr[k])] = w[e];
            }
    };
    std::vector<int32_t> pxi, pyj;
    std::vector<double> pxw, pyw;
    pad(f.xptr, f.xi, f.xw, f.nxd, h->sep_wx, pxi, pxw);
    pad(f.yptr, f.yj, f.yw, f.nyd, h->sep_wy, pyj, pyw);
    cudaError_t e = cudaMalloc(&h->d_redo, sizeof(int) * (2 + 2 * (size_t)dccm_remap::kRedoCap));
    if (e == cudaSuccess) e = cudaMemset(h->d_redo, 0, sizeof(int) * 2);
    if (e == cudaSuccess) e = upload(h->d_xi, pxi);
    if (e == cudaSuccess) e = upload(h->d_xw, pxw);
    if (e == cudaSuccess) e = upload(h->d_yj, pyj);
    if (e == cudaSuccess) e = upload(h->d_yw, pyw);
    if (e != cudaSuccess) {
        dccm_remap_destroy(h);
        return fail(DCCM_ERR_CUDA, "dccm_remap_create (separable): %s", cudaGetErrorString(e));
    }
    *out = h;
    return DCCM_OK;
}

// kind-1 handle straight from the per-row stencils of the generator (equal longitudes / axisymmetric source): the
// O(nx*ny) table is never generated.  Same limits as the detection of dccm_remap_create_lonlat (at least two
// destination columns, stencils of at most 64 entries); anything else goes through the table.
bool zonal_ok(const SepFactors &f)
{
    if (!f.zonal || f.nxd < 2 || !(f.nxs == f.nxd || f.nxs == 1)) return false;
    for (int jD = 0; jD < f.nyd; jD++)
        if (f.zptr[jD + 1] - f.zptr[jD] > 64) return false;
    return true;
}

int create_zonal(const SepFactors &f, dccm_remap **out)
{
    int rc = ensure_device();
    if (rc) return rc;
    dccm_remap *h = new dccm_remap();
    h->n_send = f.nxs * f.nys; h->n_recv = f.nxd * f.nyd;
    h->nnz = (int64_t)f.zw.size() * f.nxd;
    for (int jD = 0; jD < f.nyd; jD++) h->max_row_nnz = std::max(h->max_row_nnz, f.zptr[jD + 1] - f.zptr[jD]);
    cudaError_t e = cudaMalloc(&h->d_redo, sizeof(int) * (2 + 2 * (size_t)dccm_remap::kRedoCap));
    if (e == cudaSuccess) e = cudaMemset(h->d_redo, 0, sizeof(int) * 2);
    if (e != cudaSuccess) { dccm_remap_destroy(h); return fail(DCCM_ERR_CUDA, "dccm_remap_create (zonal): %s", cudaGetErrorString(e)); }
    rc = finish_zonal(h, f.nxs, f.nxd, f.nyd, f.zptr, f.zdi, f.zjs, f.zw);
    if (rc) { dccm_remap_destroy(h); return rc; }
    *out = h;
    return DCCM_OK;
}

// expanded-table route for the pairs the separable form does not take (zonal stencils, 2nd order)
int create_from_table(dccm_table *t, int nxs, int nys, int nxd, int nyd, dccm_remap **out)
{
    const int64_t n = dccm_table_size(t);
    std::vector<int32_t> si((size_t)n), ri((size_t)n);
    std::vector<double> cf((size_t)n);
    int rc = dccm_table_index(t, nxs, nxd, si.data(), ri.data(), cf.data());
    dccm_table_free(t);
    if (rc) return rc;
    return dccm_remap_create_lonlat(n, si.data(), ri.data(), cf.data(), nxs * nys, nxd * nyd, nxs, nxd, out);
}
}  // namespace

namespace {
// Latitude band of an operator: destination rows [j0, j1) only, source rows renumbered from src_row0 (the band's
// source buffer holds src_rows rows: own rows + halo).  The longitude factors are untouched; the latitude lists and
// the per-row stencils are cut to the band's rows -- exactly the lines of the full table that belong to them.
int band_check(int nyd, int nys, int j0, int j1, int src_row0, int src_rows)
{
    if (j0 < 0 || j1 > nyd || j0 >= j1 || src_row0 < 0 || src_rows < 1 || src_row0 + src_rows > nys)
        return fail(DCCM_ERR_ARG, "dccm_remap_create (band): rows [%d,%d) of %d, source rows [%d,%d) of %d", j0, j1, nyd,
                    src_row0, src_row0 + src_rows, nys);
    return DCCM_OK;
}

// table route of a band: the band's lines of the table, indices made local
int create_band_from_table(dccm_table *t, int nxs, int nxd, int j0, int j1, int src_row0, int src_rows, dccm_remap **out)
{
    const int64_t n = dccm_table_size(t);
    std::vector<int32_t> si((size_t)n), ri((size_t)n);
    std::vector<double> cf((size_t)n);
    int rc = dccm_table_index(t, nxs, nxd, si.data(), ri.data(), cf.data());
    dccm_table_free(t);
    if (rc) return rc;
    for (int64_t k = 0; k < n; k++) { si[k] -= src_row0 * nxs; ri[k] -= j0 * nxd; }
    return dccm_remap_create_lonlat(n, si.data(), ri.data(), cf.data(), nxs * src_rows, nxd * (j1 - j0), nxs, nxd, out);
}
}  // namespace

extern "C" int dccm_remap_create_jones99_band(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                              int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                              const double *y_LatIntWtS, const double *y_LatIntWtD,
                                              int accuracy_order, int lon_mode,
                                              int jD0, int jD1, int src_row0, int src_rows, dccm_remap **out)
{
    *out = nullptr;
    int rc = band_check(nyd, nys, jD0, jD1, src_row0, src_rows);
    if (rc) return rc;
    SepFactors f;
    rc = jones99_factors(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                         accuracy_order, lon_mode, f);
    if (rc) return rc;
    if (zonal_ok(f) || f.ok) {
        rc = slice_rows(f, jD0, jD1, src_row0, src_rows);
        if (rc) return rc;
        return zonal_ok(f) ? create_zonal(f, out) : create_separable(f, out);
    }
    dccm_table *t = nullptr;
    rc = dccm_table_gen_jones99_rows(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                                     accuracy_order, lon_mode, jD0 + 1, jD1, &t);
    if (rc) return rc;
    return create_band_from_table(t, nxs, nxd, jD0, jD1, src_row0, src_rows, out);
}

extern "C" int dccm_remap_create_bilinear_band(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                               int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                               int lon_mode, int jD0, int jD1, int src_row0, int src_rows,
                                               dccm_remap **out)
{
    *out = nullptr;
    int rc = band_check(nyr, nys, jD0, jD1, src_row0, src_rows);
    if (rc) return rc;
    SepFactors f;
    rc = bilinear_factors(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, f);
    if (rc) return rc;
    if (zonal_ok(f) || f.ok) {
        rc = slice_rows(f, jD0, jD1, src_row0, src_rows);
        if (rc) return rc;
        return zonal_ok(f) ? create_zonal(f, out) : create_separable(f, out);
    }
    dccm_table *t = nullptr;
    rc = dccm_table_gen_bilinear_rows(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, jD0 + 1, jD1, &t);
    if (rc) return rc;
    return create_band_from_table(t, nxs, nxr, jD0, jD1, src_row0, src_rows, out);
}

extern "C" int dccm_remap_create_jones99(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                         int nxd, const double *x_LonD, int nyd, const double *y_LatD,
                                         const double *y_LatIntWtS, const double *y_LatIntWtD,
                                         int accuracy_order, int lon_mode, dccm_remap **out)
{
    *out = nullptr;
    SepFactors f;
    int rc = jones99_factors(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                             accuracy_order, lon_mode, f);
    if (rc) return rc;
    if (zonal_ok(f)) return create_zonal(f, out);
    if (f.ok) return create_separable(f, out);
    dccm_table *t = nullptr;
    rc = dccm_table_gen_jones99(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, y_LatIntWtS, y_LatIntWtD,
                                accuracy_order, lon_mode, &t);
    if (rc) return rc;
    return create_from_table(t, nxs, nys, nxd, nyd, out);
}

extern "C" int dccm_remap_create_bilinear(int nxs, const double *x_LonS, int nys, const double *y_LatS,
                                          int nxr, const double *x_LonR, int nyr, const double *y_LatR,
                                          int lon_mode, dccm_remap **out)
{
    *out = nullptr;
    SepFactors f;
    int rc = bilinear_factors(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, f);
    if (rc) return rc;
    if (zonal_ok(f)) return create_zonal(f, out);         // equal longitudes: one stencil per destination row
    if (f.ok) return create_separable(f, out);
    dccm_table *t = nullptr;
    rc = dccm_table_gen_bilinear(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, &t);
    if (rc) return rc;
    return create_from_table(t, nxs, nys, nxr, nyr, out);
}

extern "C" void dccm_remap_destroy(dccm_remap *h)
{
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(g_reg_mutex);
        for (auto it = g_registry.begin(); it != g_registry.end();)
            it = (it->second == h) ? g_registry.erase(it) : std::next(it);
    }
    cudaFree(h->d_rowptr); cudaFree(h->d_col); cudaFree(h->d_w);
    cudaFree(h->d_zptr); cudaFree(h->d_zdj); cudaFree(h->d_zw); cudaFree(h->d_redo);
    cudaFree(h->d_xi); cudaFree(h->d_yj); cudaFree(h->d_xw); cudaFree(h->d_yw);
    h->send_buf.release(); h->recv_buf.release();
    delete h;
}

extern "C" int64_t dccm_remap_nnz(const dccm_remap *h) { return h ? h->nnz : -1; }
extern "C" int dccm_remap_kind(const dccm_remap *h) { return h ? h->kind : -1; }

extern "C" int dccm_remap_apply_device(dccm_remap *h, const double *d_send, int sn1,
                                       double *d_recv, int rn1, int rn2, int num_of_data, void *stream)
{
    dccm_src_seg s{d_send, d_send, d_send, 0, INT64_MAX};
    return dccm_remap_apply_seg_device(h, &s, sn1, d_recv, rn1, rn2, num_of_data, stream);
}

extern "C" int dccm_remap_apply_seg_device(dccm_remap *h, const dccm_src_seg *seg, int sn1,
                                           double *d_recv, int rn1, int rn2, int num_of_data, void *stream)
{
    NvtxRange nvtx("dccm_remap_apply_seg_device");
    if (!h || !seg) return fail(DCCM_ERR_ARG, "dccm_remap_apply: null handle");
    const SrcSeg d_send = seg_of(seg);
    if (sn1 < h->n_send || rn1 < h->n_recv)
        return fail(DCCM_ERR_ARG, "dccm_remap_apply: sn1=%d < n_send=%d or rn1=%d < n_recv=%d", sn1, h->n_send, rn1, h->n_recv);
    if (num_of_data < 0 || num_of_data > rn2)
        return fail(DCCM_ERR_ARG, "dccm_remap_apply: num_of_data=%d exceeds rn2=%d", num_of_data, rn2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // recv_data(:,:) = 0 for everything the kernel does not overwrite (ref :293)
    if (rn2 > num_of_data)
        DCCM_CUDA_TRY(cudaMemsetAsync(d_recv + (int64_t)num_of_data * rn1, 0,
                                      sizeof(double) * (size_t)(rn2 - num_of_data) * rn1, st));
    if (rn1 > h->n_recv && num_of_data > 0)
        DCCM_CUDA_TRY(cudaMemset2DAsync(d_recv + h->n_recv, sizeof(double) * (size_t)rn1, 0,
                                        sizeof(double) * (size_t)(rn1 - h->n_recv), num_of_data, st));
    if (num_of_data == 0) return DCCM_OK;
    const int gx = (h->n_recv + kThreads - 1) / kThreads;
    const bool use_seg = !(d_send.b0 <= 0 && d_send.b1 >= (int64_t)INT32_MAX);
    // field block: the smallest compiled size that takes the whole call in one pass; when rows are too few
    // to fill the machine the fields are split over grid.y instead (blocks of 2)
    static const int kFB[] = {2, 4, 5, 8, 10, 13};
    int FB = 13;
    for (int f : kFB) if (num_of_data <= f) { FB = f; break; }
    const int want = 4 * num_sms();
    if (gx < want) {          // few rows: the largest block that still leaves enough field groups to fill the machine
        const int groups = (want + gx - 1) / gx;
        FB = 2;
        for (int f : kFB) if (f <= num_of_data && (num_of_data + f - 1) / f >= groups) FB = f;
    }
    // separable operators re-derive their entries from L1-resident factors, so splitting the fields over more
    // threads costs no table traffic and buys occupancy: blocks of at most kSepFB fields, one block per grid.y
    constexpr int kSepFB = 13;
    const bool split_sep = h->kind == 2 && num_of_data > kSepFB;
    if (split_sep) { FB = 2; for (int f : kFB) if (f <= kSepFB) FB = f; }
    int nfb = (num_of_data + FB - 1) / FB;
    int gy = split_sep ? nfb : std::min(nfb, std::max(1, (want + gx - 1) / gx));
    int fields_per_y = ((nfb + gy - 1) / gy) * FB;
    gy = (num_of_data + fields_per_y - 1) / fields_per_y;
    dim3 grid(gx, gy);
#define DCCM_REMAP_LAUNCH(F, S)                                                                                   \
    do {                                                                                                          \
        if (h->kind == 2)                                                                                         \
            remap_sep_kernel<F, S><<<grid, kThreads, 0, st>>>(sep_of(h), d_send, sn1, d_recv, rn1, h->n_recv,      \
                                                             num_of_data, fields_per_y);                         \
        else if (h->kind == 1)                                                                                    \
            remap_zonal_kernel<F, S><<<grid, kThreads, 0, st>>>(h->d_zptr, h->d_zdj, h->d_zw, h->nxs, h->nxd,     \
                                                               d_send, sn1, d_recv, rn1, h->n_recv, num_of_data, \
                                                               fields_per_y);                                    \
        else                                                                                                      \
            remap_csr_kernel<F, S><<<grid, kThreads, 0, st>>>(h->d_rowptr, h->d_col, h->d_w, d_send, sn1, d_recv, \
                                                             rn1, h->n_recv, num_of_data, fields_per_y);         \
    } while (0)
#define DCCM_REMAP_FB(F) do { if (use_seg) DCCM_REMAP_LAUNCH(F, true); else DCCM_REMAP_LAUNCH(F, false); } while (0)
    switch (FB) {
    case 2: DCCM_REMAP_FB(2); break;
    case 4: DCCM_REMAP_FB(4); break;
    case 5: DCCM_REMAP_FB(5); break;
    case 8: DCCM_REMAP_FB(8); break;
    case 10: DCCM_REMAP_FB(10); break;
    default: DCCM_REMAP_FB(13); break;
    }
#undef DCCM_REMAP_FB
#undef DCCM_REMAP_LAUNCH
    DCCM_CUDA_TRY(cudaGetLastError());
    return DCCM_OK;
}

// Host form.  Large calls move their fields in groups: group g+1 on its way in (H2D stream) while the kernel runs on
// group g and group g-1 is on its way out (D2H stream) -- fields are independent, so a group is just a call with
// fewer fields on offset pointers.  The zero fill of recv_data(:, num_of_data+1:) (ref :293) is done once up front.
extern "C" int dccm_remap_apply_host(dccm_remap *h, const double *send, int sn1, int sn2,
                                     double *recv, int rn1, int rn2, int num_of_data)
{
    NvtxRange nvtx("dccm_remap_apply_host");
    if (!h) return fail(DCCM_ERR_ARG, "dccm_remap_apply: null handle");
    if (num_of_data > sn2) return fail(DCCM_ERR_ARG, "dccm_remap_apply: num_of_data=%d exceeds sn2=%d", num_of_data, sn2);
    if (num_of_data < 0 || num_of_data > rn2)
        return fail(DCCM_ERR_ARG, "dccm_remap_apply: num_of_data=%d exceeds rn2=%d", num_of_data, rn2);
    int rc = h->send_buf.reserve(sizeof(double) * (size_t)sn1 * std::max(1, num_of_data));
    if (rc) return rc;
    rc = h->recv_buf.reserve(sizeof(double) * (size_t)rn1 * std::max(1, rn2));
    if (rc) return rc;
    static thread_local cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    for (auto &ps : pipe)
        if (!ps) DCCM_CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    cudaStream_t sin = pipe[0], sk = pipe[1], sout = pipe[2];
    const char *env = getenv("DCCM_HOST_CHUNKS");
    const bool big = (size_t)(sn1 + rn1) * (size_t)std::max(1, num_of_data) >= ((size_t)1 << 22);
    int ngroup = env ? std::max(1, atoi(env)) : (big ? 4 : 1);
    ngroup = std::max(1, std::min(ngroup, num_of_data));
    double *d_send = h->send_buf.as<double>(), *d_recv = h->recv_buf.as<double>();
    // rows the kernel does not write: recv_data(:,:) = 0 (ref :293) -- set on the host, nothing to copy back
    if (rn2 > num_of_data) memset(recv + (size_t)num_of_data * rn1, 0, sizeof(double) * (size_t)(rn2 - num_of_data) * rn1);
    std::vector<cudaEvent_t> ev(2 * (size_t)ngroup, nullptr);
    struct EvGuard { std::vector<cudaEvent_t> &e; ~EvGuard() { for (auto x : e) if (x) cudaEventDestroy(x); } } guard{ev};
    for (auto &e : ev) DCCM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const int per = num_of_data > 0 ? (num_of_data + ngroup - 1) / ngroup : 0;
    for (int g = 0; g < ngroup && num_of_data > 0; g++) {
        const int f0 = g * per, f1 = std::min(num_of_data, f0 + per);
        if (f0 >= f1) continue;
        const size_t so = (size_t)f0 * sn1, ro = (size_t)f0 * rn1;
        DCCM_CUDA_TRY(cudaMemcpyAsync(d_send + so, send + so, sizeof(double) * (size_t)sn1 * (f1 - f0), cudaMemcpyHostToDevice, sin));
        DCCM_CUDA_TRY(cudaEventRecord(ev[2 * g], sin));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sk, ev[2 * g], 0));
        rc = dccm_remap_apply_device(h, d_send + so, sn1, d_recv + ro, rn1, f1 - f0, f1 - f0, sk);
        if (rc) return rc;
        DCCM_CUDA_TRY(cudaEventRecord(ev[2 * g + 1], sk));
        DCCM_CUDA_TRY(cudaStreamWaitEvent(sout, ev[2 * g + 1], 0));
        DCCM_CUDA_TRY(cudaMemcpyAsync(recv + ro, d_recv + ro, sizeof(double) * (size_t)rn1 * (f1 - f0), cudaMemcpyDeviceToHost, sout));
    }
    DCCM_CUDA_TRY(cudaStreamSynchronize(sout));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sk));
    DCCM_CUDA_TRY(cudaStreamSynchronize(sin));
    return DCCM_OK;
}

extern "C" int dccm_interp_register(int recv_model, int send_model, int mapping_tag, dccm_remap *h)
{
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    g_registry[std::make_tuple(recv_model, send_model, mapping_tag)] = h;
    return DCCM_OK;
}

extern "C" int dccm_interpolate_data(int recv_model, int send_model, int mapping_tag,
                                     int sn1, int sn2, const double *send_data,
                                     int rn1, int rn2, double *recv_data, int num_of_data)
{
    dccm_remap *h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_reg_mutex);
        auto it = g_registry.find(std::make_tuple(recv_model, send_model, mapping_tag));
        if (it != g_registry.end()) h = it->second;
    }
    if (!h)
        return fail(DCCM_ERR_ARG, "interpolate_data: no operation index registered for (recv=%d, send=%d, tag=%d)",
                    recv_model, send_model, mapping_tag);
    return dccm_remap_apply_host(h, send_data, sn1, sn2, recv_data, rn1, rn2, num_of_data);
}

// ---------------------------------------------------------------- the coupler's callback, without a Fortran shim

extern "C" int dccm_interp_set_model_name(int model_id, const char *name)
{
    if (!name || !*name || model_id < 1) return fail(DCCM_ERR_ARG, "dccm_interp_set_model_name: bad arguments");
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    g_model_ids[fortran_name(name, (int64_t)strlen(name))] = model_id;
    return DCCM_OK;
}

extern "C" int dccm_interpolate_data_named(const char *recv_model, int64_t recv_len,
                                           const char *send_model, int64_t send_len, int mapping_tag,
                                           int sn1, int sn2, const double *send_data,
                                           int rn1, int rn2, double *recv_data, int num_of_data)
{
    const std::string r = fortran_name(recv_model, recv_len), s = fortran_name(send_model, send_len);
    int ir = 0, is = 0;
    {
        std::lock_guard<std::mutex> lk(g_reg_mutex);
        auto a = g_model_ids.find(r), b = g_model_ids.find(s);
        if (a != g_model_ids.end()) ir = a->second;
        if (b != g_model_ids.end()) is = b->second;
    }
    if (!ir || !is)
        return fail(DCCM_ERR_ARG, "interpolate_data: unknown component name '%s' (dccm_interp_set_model_name)",
                    (!ir ? r : s).c_str());
    return dccm_interpolate_data(ir, is, mapping_tag, sn1, sn2, send_data, rn1, rn2, recv_data, num_of_data);
}

namespace {
void default_f77_error(const char *msg)
{
    fprintf(stderr, "interpolate_data: %s\n", msg);
    fflush(stderr);
    abort();       // the reference leaves through jcup_error, which aborts the WHOLE run: under MPI a plain exit(1) of one
                   // rank would leave the others hanging in Jcup's collectives; hosts install dccm_f77_set_error_handler
}
void (*g_f77_error)(const char *) = default_f77_error;
}  // namespace

extern "C" void dccm_f77_set_error_handler(void (*handler)(const char *))
{
    g_f77_error = handler ? handler : default_f77_error;
}

// The external procedure itself under its Fortran link name (gfortran / ifort / nvfortran on x86-64 and aarch64:
// lower case + underscore, every argument by reference, the CHARACTER lengths appended by value).  Lengths are
// size_t with gfortran >= 8 and 32-bit int before that; only the low 32 bits are looked at, so both work.
extern "C" void interpolate_data_(const char *recv_model, const char *send_model, const int32_t *mapping_tag,
                                  const int32_t *sn1, const int32_t *sn2, const double *send_data,
                                  const int32_t *rn1, const int32_t *rn2, double *recv_data,
                                  const int32_t *num_of_data, const int32_t *tn, const int32_t *exchange_tag,
                                  size_t recv_model_len, size_t send_model_len)
{
    (void)tn; (void)exchange_tag;                  // unused by the reference as well (ref common/interpolate_data.f90:15)
    int rc = dccm_interpolate_data_named(recv_model, (int64_t)(recv_model_len & 0xffffffffu),
                                         send_model, (int64_t)(send_model_len & 0xffffffffu), *mapping_tag,
                                         *sn1, *sn2, send_data, *rn1, *rn2, recv_data, *num_of_data);
    if (rc != DCCM_OK) g_f77_error(dccm_last_error());
}
