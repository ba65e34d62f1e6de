// bcsr.cu -- y = B x for 4x4 block CSR with row-major blocks (nsk_bcsr4_* of nsk.h).
//
// Replaces SpMV_BCSR / _OPT / _FMA / _AVX2 (reference mpk/SpMV.cpp:90-219) on the container
// bcsr4x4_matrix (mpk/SpMV.h:26-33) as built by generate_BCSR4 (mpk/utils.cpp:45-95: block columns in
// first-appearance order, explicit zeros inside blocks).  Index traffic drops from 4 B/nnz (CSR) to
// 0.25 B/nnz; the value stream (8 B/nnz incl. explicit zeros) is what HBM sees.
//
// One thread per SCALAR row: thread (bi, i) walks the blocks of block row bi in storage order and
// accumulates sum_j blk[4i+j] * x[4bj+j] with j = 0..3 innermost -- exactly the (block, j) order of
// SpMV_BCSR_FMA, so the exact modes are bit-identical to the reference (lane i of the AVX2 variant
// performs the same chain).  The four threads of a block row read one 128-byte block per step (one
// fully used line); x[4bj..4bj+3] is a 32-byte sector shared by the four threads through L1.
#include "nsk_internal.h"
#include "stream_common.cuh"

// The value stream is read exactly once: no L1 allocation, so L1 keeps the x sectors the four threads of a block row share.
__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Loads of BATCH consecutive blocks (block column, two 16-byte halves of the thread's block row, the 32-byte x sector)
// are all issued before the first dependent fma, so one thread keeps 4 x BATCH 16-byte loads in flight instead of 4: the
// chain itself stays strictly in (block, j) order.
template <bool MULADD, int BATCH>
__global__ void __launch_bounds__(256) spmv_bcsr4_kernel(int nbrows, const int *__restrict__ ptrow,
                                                         const int *__restrict__ indcol,
                                                         const double *__restrict__ coef,
                                                         const double *__restrict__ x, double *__restrict__ y)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int bi = row >> 2, i = row & 3;
    if (bi >= nbrows) return;
    const int p = ptrow[bi], q = ptrow[bi + 1];
    double acc = 0.0;
    int ia = p;
    for (; ia + BATCH <= q; ia += BATCH) {
        int bj[BATCH];
        double2 a01[BATCH], a23[BATCH], x01[BATCH], x23[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) bj[u] = __ldg(indcol + ia + u);
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const double2 *blk = reinterpret_cast<const double2 *>(coef + 16 * (size_t)(ia + u) + 4 * i);
            a01[u] = ld_stream(blk);
            a23[u] = ld_stream(blk + 1);
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const double2 *xv = reinterpret_cast<const double2 *>(x + 4 * (size_t)bj[u]);
            x01[u] = __ldg(xv);
            x23[u] = __ldg(xv + 1);
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            acc = row_op<MULADD>(a01[u].x, x01[u].x, acc);
            acc = row_op<MULADD>(a01[u].y, x01[u].y, acc);
            acc = row_op<MULADD>(a23[u].x, x23[u].x, acc);
            acc = row_op<MULADD>(a23[u].y, x23[u].y, acc);
        }
    }
    for (; ia < q; ia++) {
        const int bj = __ldg(indcol + ia);
        const double2 *blk = reinterpret_cast<const double2 *>(coef + 16 * (size_t)ia + 4 * i);
        const double2 *xv = reinterpret_cast<const double2 *>(x + 4 * (size_t)bj);
        const double2 a01 = ld_stream(blk), a23 = ld_stream(blk + 1);
        const double2 x01 = __ldg(xv), x23 = __ldg(xv + 1);
        acc = row_op<MULADD>(a01.x, x01.x, acc);
        acc = row_op<MULADD>(a01.y, x01.y, acc);
        acc = row_op<MULADD>(a23.x, x23.x, acc);
        acc = row_op<MULADD>(a23.y, x23.y, acc);
    }
    y[row] = acc;
}

NSK_API int nsk_bcsr4_create(nsk_ctx_t ctx, int nbrows, int64_t nblocks, const int *ptrow, const int *indcol,
                             const double *coef, nsk_bcsr4_t *out)
{
    if (!ctx || !out) return NSK_ERR_INVALID;
    *out = nullptr;
    NSK_REQUIRE(ctx, nbrows >= 0 && nblocks >= 0 && ptrow, "bad sizes");
    NSK_REQUIRE(ctx, ptrow[0] == 0 && (int64_t)ptrow[nbrows] == nblocks, "ptrow[nbrows] must equal nblocks");
    NSK_REQUIRE(ctx, nblocks == 0 || (indcol && coef), "indcol/coef null");
    for (int b = 0; b < nbrows; b++) NSK_REQUIRE(ctx, ptrow[b + 1] >= ptrow[b], "ptrow decreases");
    for (int64_t e = 0; e < nblocks; e++)
        NSK_REQUIRE(ctx, indcol[e] >= 0 && indcol[e] < nbrows, "block column out of range");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    nsk_bcsr4_s *B = new nsk_bcsr4_s();
    B->ctx = ctx;
    B->nbrows = nbrows;
    B->nblocks = nblocks;
    if (cudaMalloc(&B->d_ptrow, sizeof(int) * ((size_t)nbrows + 1)) != cudaSuccess ||
        cudaMalloc(&B->d_indcol, sizeof(int) * (size_t)std::max<int64_t>(nblocks, 1)) != cudaSuccess ||
        cudaMalloc(&B->d_coef, sizeof(double) * 16 * (size_t)std::max<int64_t>(nblocks, 1)) != cudaSuccess) {
        nsk_set_error(ctx, "cudaMalloc of the block operator failed");
        nsk_bcsr4_destroy(B);
        return NSK_ERR_ALLOC;
    }
    NSK_CUDA(ctx, cudaMemcpy(B->d_ptrow, ptrow, sizeof(int) * ((size_t)nbrows + 1), cudaMemcpyHostToDevice));
    if (nblocks) {
        NSK_CUDA(ctx, cudaMemcpy(B->d_indcol, indcol, sizeof(int) * (size_t)nblocks, cudaMemcpyHostToDevice));
        NSK_CUDA(ctx, cudaMemcpy(B->d_coef, coef, sizeof(double) * 16 * (size_t)nblocks, cudaMemcpyHostToDevice));
    }
    *out = B;
    return NSK_OK;
}

NSK_API int nsk_bcsr4_destroy(nsk_bcsr4_t B)
{
    if (!B) return NSK_OK;
    cudaSetDevice(B->ctx->device);
    cudaStreamSynchronize(B->ctx->stream);
    if (B->d_ptrow) cudaFree(B->d_ptrow);
    if (B->d_indcol) cudaFree(B->d_indcol);
    if (B->d_coef) cudaFree(B->d_coef);
    delete B;
    return NSK_OK;
}

NSK_API int nsk_spmv_bcsr4(nsk_bcsr4_t B, const double *x, double *y, nsk_mode mode, nsk_where where)
{
    if (!B) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = B->ctx;
    NSK_REQUIRE(ctx, x && y, "x or y is null");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nb = sizeof(double) * 4 * (size_t)B->nbrows;
    const double *dx = x;
    double *dy = y;
    if (where == NSK_HOST) {
        void *vx, *vy;
        NSK_TRY(nsk_stage(ctx, 0, nb, &vx));
        NSK_TRY(nsk_stage(ctx, 1, nb, &vy));
        NSK_CUDA(ctx, cudaMemcpyAsync(vx, x, nb, cudaMemcpyHostToDevice, ctx->stream));
        dx = (const double *)vx;
        dy = (double *)vy;
    }
    NSK_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(dx) & 15) == 0, "x must be 16-byte aligned");
    const int rows = 4 * B->nbrows;
    if (rows > 0) {
        const int blocks = (rows + 255) / 256;
        const int batch = ctx->opt.bcsr_batch > 0 ? (int)ctx->opt.bcsr_batch : 4;
#define NSK_BCSR_LAUNCH(MA, BT) \
    spmv_bcsr4_kernel<MA, BT><<<blocks, 256, 0, ctx->stream>>>(B->nbrows, B->d_ptrow, B->d_indcol, B->d_coef, dx, dy)
        const bool ma = mode == NSK_EXACT_MULADD;
        if (batch >= 4) { if (ma) NSK_BCSR_LAUNCH(true, 4); else NSK_BCSR_LAUNCH(false, 4); }
        else if (batch >= 2) { if (ma) NSK_BCSR_LAUNCH(true, 2); else NSK_BCSR_LAUNCH(false, 2); }
        else { if (ma) NSK_BCSR_LAUNCH(true, 1); else NSK_BCSR_LAUNCH(false, 1); }
#undef NSK_BCSR_LAUNCH
        ctx->launches++;
        NSK_CUDA(ctx, cudaGetLastError());
    }
    if (where == NSK_HOST) {
        NSK_CUDA(ctx, cudaMemcpyAsync(y, dy, nb, cudaMemcpyDeviceToHost, ctx->stream));
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NSK_OK;
}
