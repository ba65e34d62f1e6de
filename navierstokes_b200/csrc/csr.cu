// csr.cu -- CSR operator object and the public SpMV / matrix-powers entry points of nsk.h.
//
// nsk_csr_create   <- struct csrmatrix                      (reference mpk/SpMV.h:18-24)
// nsk_spmv         <- SpMV_CSR / _OPT / _FMA / _AVX2        (reference mpk/SpMV.cpp:6-85)
// nsk_mpk          <- SpM2V_CSR* / SpM3V / SpM4V            (reference mpk/SpM2V.cpp:80-332,
//                                                            mpk/SpMVmulti0.cpp:132-221)
#include <map>
#include <mutex>

#include <algorithm>
#include <thread>
#include <vector>

#include "nsk_internal.h"

// host copy of ptrow per operator (needed to (re)build tilings for other tile geometries)
static std::map<nsk_csr_t, std::vector<int>> g_host_ptrow;
static std::mutex g_host_ptrow_mu;

std::vector<int> &nsk_csr_host_ptrow(nsk_csr_t A)
{
    std::lock_guard<std::mutex> lk(g_host_ptrow_mu);
    return g_host_ptrow[A];
}

int nsk_mpk_device(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode);  // mpk.cu
int nsk_mpk_device2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                    double *const *d_levels2, nsk_mode mode);  // mpk.cu
void nsk_wave_set_block_extents(nsk_csr_t A, const int *ptrow, const int *indcol);  // wave_deps.cu
void nsk_wave_free(nsk_csr_t A);

NSK_API int nsk_csr_create(nsk_ctx_t ctx, int n, int n_cols, int64_t nnz, const int *ptrow,
                           const int *indcol, const double *coef, nsk_csr_t *out)
{
    if (!ctx || !out) return NSK_ERR_INVALID;
    *out = nullptr;
    NSK_REQUIRE(ctx, n >= 0 && n_cols >= 0 && nnz >= 0, "negative size");
    NSK_REQUIRE(ctx, nnz < (int64_t)2147483647, "nnz must fit the reference's int (mpk/SpMV.h:20)");
    NSK_REQUIRE(ctx, ptrow != nullptr, "ptrow is null");
    NSK_REQUIRE(ctx, nnz == 0 || (indcol && coef), "indcol/coef null");
    NSK_REQUIRE(ctx, ptrow[0] == 0, "ptrow[0] must be 0");
    NSK_REQUIRE(ctx, (int64_t)ptrow[n] == nnz, "ptrow[n] must equal nnz");
    int max_row = 0;
    for (int i = 0; i < n; i++) {
        int len = ptrow[i + 1] - ptrow[i];
        if (len < 0) {
            nsk_set_error(ctx, "ptrow decreases at row %d", i);
            return NSK_ERR_INVALID;
        }
        if (len > max_row) max_row = len;
    }
    {
        // column range check, threaded (117 M entries for a 256^3 operator: part of the one-off upload cost bench.py reports)
        const int nth = nnz > (int64_t)1 << 22 ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
        std::vector<int64_t> first_bad((size_t)nth, -1);
        auto scan = [&](int t) {
            const int64_t e0 = nnz * t / nth, e1 = nnz * (t + 1) / nth;
            for (int64_t e = e0; e < e1; e++)
                if (indcol[e] < 0 || indcol[e] >= n_cols) { first_bad[(size_t)t] = e; return; }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nth; t++) th.emplace_back(scan, t);
        scan(0);
        for (auto &x : th) x.join();
        for (int t = 0; t < nth; t++)
            if (first_bad[(size_t)t] >= 0) {
                const int64_t e = first_bad[(size_t)t];
                nsk_set_error(ctx, "column index %d out of range [0,%d) at entry %lld", indcol[e], n_cols, (long long)e);
                return NSK_ERR_INVALID;
            }
    }
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    nsk_csr_s *A = new nsk_csr_s();
    A->ctx = ctx;
    A->n = n;
    A->n_cols = n_cols;
    A->nnz = nnz;
    A->max_row = max_row;
    A->mean_row = n ? (double)nnz / (double)n : 0.0;
    // +64 B: the bulk copies round slices up to 16-byte granules and may read past the end
    const size_t pad = 64;
    cudaError_t e1 = cudaMalloc(&A->d_ptrow, sizeof(int) * ((size_t)n + 1) + pad);
    cudaError_t e2 = cudaMalloc(&A->d_indcol, sizeof(int) * (size_t)nnz + pad);
    cudaError_t e3 = cudaMalloc(&A->d_coef, sizeof(double) * (size_t)nnz + pad);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        nsk_set_error(ctx, "cudaMalloc of the operator (%lld nnz) failed", (long long)nnz);
        nsk_csr_destroy(A);
        return NSK_ERR_ALLOC;
    }
    NSK_CUDA(ctx, cudaMemset((char *)A->d_ptrow + sizeof(int) * ((size_t)n + 1), 0, pad));
    NSK_CUDA(ctx, cudaMemset((char *)A->d_indcol + sizeof(int) * (size_t)nnz, 0, pad));
    NSK_CUDA(ctx, cudaMemset((char *)A->d_coef + sizeof(double) * (size_t)nnz, 0, pad));
    NSK_CUDA(ctx, cudaMemcpy(A->d_ptrow, ptrow, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        NSK_CUDA(ctx, cudaMemcpy(A->d_indcol, indcol, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        NSK_CUDA(ctx, cudaMemcpy(A->d_coef, coef, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    nsk_csr_host_ptrow(A).assign(ptrow, ptrow + n + 1);
    nsk_wave_set_block_extents(A, ptrow, indcol);
    *out = A;  // tile tables are built on first use (a distributed slab registers its breaks first)
    return NSK_OK;
}

void nsk_dist_free(nsk_csr_t A);  // dist.cu

NSK_API int nsk_csr_destroy(nsk_csr_t A)
{
    if (!A) return NSK_OK;
    nsk_ctx_t ctx = A->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nsk_dist_free(A);
    nsk_wave_free(A);
    nsk_packed_free(A);
    nsk_sell_free(A);
    if (A->d_ptrow) cudaFree(A->d_ptrow);
    if (A->d_indcol) cudaFree(A->d_indcol);
    if (A->d_coef) cudaFree(A->d_coef);
    nsk_free_tilings(A);
    if (A->d_flags) cudaFree(A->d_flags);
    {
        std::lock_guard<std::mutex> lk(g_host_ptrow_mu);
        g_host_ptrow.erase(A);
    }
    delete A;
    return NSK_OK;
}

NSK_API int nsk_csr_shape(nsk_csr_t A, int *n, int *n_cols, int64_t *nnz)
{
    if (!A) return NSK_ERR_INVALID;
    if (n) *n = A->n;
    if (n_cols) *n_cols = A->n_cols;
    if (nnz) *nnz = A->nnz;
    return NSK_OK;
}

NSK_API int64_t nsk_csr_spmv_bytes(nsk_csr_t A)
{
    if (!A) return 0;
    return 12 * A->nnz + 4 * ((int64_t)A->n + 1) + 16 * (int64_t)A->n;
}

NSK_API int64_t nsk_csr_mpk_bytes(nsk_csr_t A, int k)
{
    if (!A) return 0;
    return 12 * A->nnz + 4 * ((int64_t)A->n + 1) + 8 * (int64_t)A->n + 8 * (int64_t)A->n * k;
}

NSK_API int64_t nsk_csr_packed_bytes(nsk_csr_t A)
{
    if (!A) return 0;
    cudaSetDevice(A->ctx->device);
    return (int64_t)nsk_packed_bytes(A);
}

NSK_API int64_t nsk_csr_tile_bytes(nsk_csr_t A)
{
    if (!A) return 0;
    cudaSetDevice(A->ctx->device);
    return (int64_t)nsk_sell_bytes(A);
}

int nsk_halo_exchange_dev(nsk_csr_t A, double *xlocal, int depth, bool allow_push);
int nsk_halo_release_dev(nsk_csr_t A, const double *xlocal, int depth);  // dist.cu

NSK_API int nsk_spmv(nsk_csr_t A, const double *x, double *y, nsk_mode mode, nsk_where where)
{
    if (!A) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    NSK_REQUIRE(ctx, x && y, "x or y is null");
    NSK_REQUIRE(ctx, mode == NSK_EXACT_FMA || mode == NSK_EXACT_MULADD || mode == NSK_FAST, "bad mode");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n_out = nsk_csr_owned_rows(A);  // distributed: only the owned rows are produced
    nsk_spmv_args a;
    a.row_begin = 0;
    a.row_end = n_out;
    a.mode = mode;
    if (where == NSK_DEVICE) {
        if (A->dist) NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(x), 1, true));
        a.x = x;
        a.y = y;
        NSK_TRY(nsk_launch_spmv(A, a));
        return A->dist ? nsk_halo_release_dev(A, x, 1) : NSK_OK;
    }
    // host pointers: x holds the owned part (all of x for a single-GPU operator)
    const size_t n_in = A->dist ? (size_t)n_out : (size_t)A->n_cols;
    void *dx, *dy;
    NSK_TRY(nsk_stage(ctx, 0, sizeof(double) * (size_t)A->n_cols, &dx));
    NSK_TRY(nsk_stage(ctx, 1, sizeof(double) * (size_t)A->n, &dy));
    NSK_CUDA(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * n_in, cudaMemcpyHostToDevice, ctx->stream));
    if (A->dist) NSK_TRY(nsk_halo_exchange_dev(A, (double *)dx, 1, false));
    a.x = (const double *)dx;
    a.y = (double *)dy;
    NSK_TRY(nsk_launch_spmv(A, a));
    NSK_CUDA(ctx, cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)n_out, cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NSK_OK;
}

NSK_API int nsk_mpk(nsk_csr_t A, int k, const double *x, double *const *levels, nsk_mode mode,
                    nsk_where where)
{
    if (!A) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    NSK_REQUIRE(ctx, k >= 1 && k <= NSK_MAX_K, "k out of range");
    NSK_REQUIRE(ctx, x && levels, "x or levels is null");
    NSK_REQUIRE(ctx, A->dist != nullptr || A->n == A->n_cols, "matrix powers need a square operator");
    for (int l = 0; l < k; l++) NSK_REQUIRE(ctx, levels[l] != nullptr, "a level pointer is null");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (where == NSK_DEVICE) return nsk_mpk_device(A, k, x, levels, mode);

    // host pointers: owned parts in, owned parts out; level scratch is one local vector per level
    const int n_out = nsk_csr_owned_rows(A);
    const size_t nb = sizeof(double) * (size_t)n_out;
    const size_t ld = ((size_t)A->n_cols + 1) & ~(size_t)1;  // even: every level starts 16-byte aligned
    void *dx, *dl;
    NSK_TRY(nsk_stage(ctx, 0, sizeof(double) * ld, &dx));
    NSK_TRY(nsk_stage(ctx, 1, sizeof(double) * ld * (size_t)k, &dl));
    NSK_CUDA(ctx, cudaMemcpyAsync(dx, x, nb, cudaMemcpyHostToDevice, ctx->stream));
    double *dlev[NSK_MAX_K];
    for (int l = 0; l < k; l++) dlev[l] = (double *)dl + (size_t)l * ld;
    if (A->dist == nullptr && k > 1 && ctx->opt.mpk_kernel == 0 && ctx->opt.host_overlap) {
        // Host-pointer call on one GPU: PCIe is the bound (k vectors out at ~57 GB/s dwarf the products), so the levels are
        // produced one product at a time and each is copied out on a second stream while the next ones are computed --
        // the first copy starts after one product instead of after all k (measured on 256^3, k = 4: 12.5 -> 12.0 ms).
        if (!ctx->copy_stream) NSK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        const double *src = (const double *)dx;
        for (int l = 0; l < k; l++) {
            if (!ctx->copy_event[l]) NSK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_event[l], cudaEventDisableTiming));
            nsk_spmv_args a;
            a.x = src;
            a.y = dlev[l];
            a.row_begin = 0;
            a.row_end = A->n;
            a.mode = mode;
            NSK_TRY(nsk_launch_spmv(A, a));
            NSK_CUDA(ctx, cudaEventRecord(ctx->copy_event[l], ctx->stream));
            NSK_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event[l], 0));
            NSK_CUDA(ctx, cudaMemcpyAsync(levels[l], dlev[l], nb, cudaMemcpyDeviceToHost, ctx->copy_stream));
            src = dlev[l];
        }
        ctx->last_mpk = 1;
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        return NSK_OK;
    }
    NSK_TRY(nsk_mpk_device(A, k, (const double *)dx, dlev, mode));
    for (int l = 0; l < k; l++)
        NSK_CUDA(ctx, cudaMemcpyAsync(levels[l], dlev[l], nb, cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NSK_OK;
}

// Several right-hand sides (s-step / block Krylov bases; the reference's counterpart is the dense-block product
// MatMatMult_SeqBAIJ_4_AVX2(A, X, Y, s_step), src/kernels/spmm_avx2.c:7-109, and its monomial basis builder
// BuildKrylovBasis_AVX2 :112-168): levels[v * k + l] = A^(l+1) xs[v].  Device-resident vectors are taken two at a
// time through the two-vector fused kernel (each tile of the operator is streamed once per pair); host vectors go
// one by one (PCIe dominates there).
NSK_API int nsk_mpk_multi(nsk_csr_t A, int k, int nvec, const double *const *xs, double *const *levels, nsk_mode mode,
                          nsk_where where)
{
    if (!A) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    NSK_REQUIRE(ctx, k >= 1 && k <= NSK_MAX_K && nvec >= 1, "k or nvec out of range");
    NSK_REQUIRE(ctx, xs && levels, "xs or levels is null");
    NSK_REQUIRE(ctx, A->dist != nullptr || A->n == A->n_cols, "matrix powers need a square operator");
    for (int v = 0; v < nvec; v++) {
        NSK_REQUIRE(ctx, xs[v] != nullptr, "an input vector is null");
        for (int l = 0; l < k; l++) NSK_REQUIRE(ctx, levels[(size_t)v * k + l] != nullptr, "a level pointer is null");
    }
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    int v = 0;
    if (where == NSK_DEVICE)
        for (; v + 1 < nvec; v += 2)
            NSK_TRY(nsk_mpk_device2(A, k, xs[v], levels + (size_t)v * k, xs[v + 1], levels + (size_t)(v + 1) * k, mode));
    for (; v < nvec; v++) NSK_TRY(nsk_mpk(A, k, xs[v], levels + (size_t)v * k, mode, where));
    return NSK_OK;
}
