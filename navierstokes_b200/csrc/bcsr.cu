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
    B->h_ptrow.assign(ptrow, ptrow + nbrows + 1);
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
    if (B->expanded) nsk_csr_destroy(B->expanded);
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
        // blocks whose loads are issued together: 2 by default (530 k rows: 0.0472 ms with 2 or 4, 0.0503 with 1; 1.56 M rows:
        // 0.1209 / 0.1225 / 0.1270 ms = 0.986 / 0.973 / 0.939 of the copy peak; one 256-bit load per block row instead of
        // two 128-bit halves was slower: 0.0513 / 0.1312 ms -- profiles/r02_final_configs_c1.txt)
        const int batch = ctx->opt.bcsr_batch > 0 ? (int)ctx->opt.bcsr_batch : 2;
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


// -----------------------------------------------------------------------------------------------
// Y = B X for s dense right-hand sides: the s-step basis product of the reference,
// MatMatMult_SeqBAIJ_4_AVX2 (src/kernels/spmm_avx2.c:7-109), and its caller BuildKrylovBasis_AVX2 (:112-168).
//
// Arithmetic per output entry (row r of block row bi, column k), exactly as the reference groups it: per block
// acc = fma(a[r][3], x3, fma(a[r][2], x2, fma(a[r][1], x1, fma(a[r][0], x0, 0)))), then sum = sum + acc (rounded add),
// blocks in storage order.  The reference keeps the same scalar in all four lanes of a __m256d and finishes with a
// horizontal sum over the lanes (:93-99), i.e. it returns 4 * sum; this kernel stores the product itself (sum) -- the
// factor is exact in binary floating point and tests/ compare 4 * Y with the literal restatement bit for bit.
//
// One thread per scalar row and pass of NC <= 4 columns (the reference's column groups): a block row of B (32 bytes per
// thread, one 128-byte line per four threads) is loaded once per pass and used for NC columns.
template <int NC>
__global__ void __launch_bounds__(256) spmm_bcsr4_kernel(int nbrows, const int *__restrict__ ptrow, const int *__restrict__ indcol,
                                                         const double *__restrict__ coef, const double *__restrict__ X,
                                                         long long ldx, double *__restrict__ Y, long long ldy)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int bi = row >> 2, i = row & 3;
    if (bi >= nbrows) return;
    const int p = ptrow[bi], q = ptrow[bi + 1];
    double sum[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) sum[k] = 0.0;
    for (int ia = p; ia < q; ia += 2) {
        const bool two = ia + 1 < q;
        const int bj0 = __ldg(indcol + ia), bj1 = two ? __ldg(indcol + ia + 1) : bj0;
        const double2 *blk0 = reinterpret_cast<const double2 *>(coef + 16 * (size_t)ia + 4 * i);
        const double2 *blk1 = reinterpret_cast<const double2 *>(coef + 16 * (size_t)(two ? ia + 1 : ia) + 4 * i);
        const double2 a0 = ld_stream(blk0), a1 = ld_stream(blk0 + 1), b0 = ld_stream(blk1), b1 = ld_stream(blk1 + 1);
        double2 x0[NC], x1[NC], z0[NC], z1[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const double2 *xa = reinterpret_cast<const double2 *>(X + (size_t)k * ldx + 4 * (size_t)bj0);
            const double2 *xb = reinterpret_cast<const double2 *>(X + (size_t)k * ldx + 4 * (size_t)bj1);
            x0[k] = __ldg(xa);
            x1[k] = __ldg(xa + 1);
            z0[k] = __ldg(xb);
            z1[k] = __ldg(xb + 1);
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
            double acc = __fma_rn(a0.x, x0[k].x, 0.0);
            acc = __fma_rn(a0.y, x0[k].y, acc);
            acc = __fma_rn(a1.x, x1[k].x, acc);
            acc = __fma_rn(a1.y, x1[k].y, acc);
            sum[k] = __dadd_rn(sum[k], acc);
            if (two) {
                double acc2 = __fma_rn(b0.x, z0[k].x, 0.0);
                acc2 = __fma_rn(b0.y, z0[k].y, acc2);
                acc2 = __fma_rn(b1.x, z1[k].x, acc2);
                acc2 = __fma_rn(b1.y, z1[k].y, acc2);
                sum[k] = __dadd_rn(sum[k], acc2);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NC; k++) Y[(size_t)k * ldy + row] = sum[k];
}

static int spmm_bcsr4_device(nsk_bcsr4_t B, int s, const double *dX, long long ldx, double *dY, long long ldy)
{
    nsk_ctx_t ctx = B->ctx;
    const int rows = 4 * B->nbrows;
    if (rows == 0) return NSK_OK;
    const int blocks = (rows + 255) / 256;
    for (int k0 = 0; k0 < s; k0 += 4) {
        const int nc = std::min(4, s - k0);
        const double *x = dX + (size_t)k0 * ldx;
        double *y = dY + (size_t)k0 * ldy;
        switch (nc) {
        case 1: spmm_bcsr4_kernel<1><<<blocks, 256, 0, ctx->stream>>>(B->nbrows, B->d_ptrow, B->d_indcol, B->d_coef, x, ldx, y, ldy); break;
        case 2: spmm_bcsr4_kernel<2><<<blocks, 256, 0, ctx->stream>>>(B->nbrows, B->d_ptrow, B->d_indcol, B->d_coef, x, ldx, y, ldy); break;
        case 3: spmm_bcsr4_kernel<3><<<blocks, 256, 0, ctx->stream>>>(B->nbrows, B->d_ptrow, B->d_indcol, B->d_coef, x, ldx, y, ldy); break;
        default: spmm_bcsr4_kernel<4><<<blocks, 256, 0, ctx->stream>>>(B->nbrows, B->d_ptrow, B->d_indcol, B->d_coef, x, ldx, y, ldy); break;
        }
        ctx->launches++;
        NSK_CUDA(ctx, cudaGetLastError());
    }
    return NSK_OK;
}

NSK_API int nsk_spmm_bcsr4(nsk_bcsr4_t B, int s, const double *X, int64_t ldx, double *Y, int64_t ldy, nsk_where where)
{
    if (!B) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = B->ctx;
    const int64_t n = 4 * (int64_t)B->nbrows;
    NSK_REQUIRE(ctx, s >= 0 && (s == 0 || (X && Y)), "X or Y is null");
    NSK_REQUIRE(ctx, ldx >= n && ldy >= n, "leading dimensions must be at least 4 * nbrows");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (s == 0 || n == 0) return NSK_OK;
    if (where == NSK_DEVICE) {
        NSK_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx & 1) == 0, "X columns must be 16-byte aligned");
        return spmm_bcsr4_device(B, s, X, ldx, Y, ldy);
    }
    const int64_t ld = (n + 1) & ~(int64_t)1;  // device columns 16-byte aligned
    void *vx, *vy;
    NSK_TRY(nsk_stage(ctx, 0, sizeof(double) * (size_t)ld * s, &vx));
    NSK_TRY(nsk_stage(ctx, 1, sizeof(double) * (size_t)ld * s, &vy));
    NSK_CUDA(ctx, cudaMemcpy2DAsync(vx, sizeof(double) * ld, X, sizeof(double) * ldx, sizeof(double) * n, s, cudaMemcpyHostToDevice, ctx->stream));
    NSK_TRY(spmm_bcsr4_device(B, s, (const double *)vx, ld, (double *)vy, ld));
    NSK_CUDA(ctx, cudaMemcpy2DAsync(Y, sizeof(double) * ldy, vy, sizeof(double) * ld, sizeof(double) * n, s, cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NSK_OK;
}

// V(:, 0) = v0, V(:, j + 1) = B V(:, j), j = 0 .. s-1: the monomial basis of BuildKrylovBasis_AVX2 (spmm_avx2.c:112-168),
// one single-column product of the kernel above per vector (the reference calls MatMatMult_SeqBAIJ_4_AVX2(A, X_k, Y_k, 1)).
NSK_API int nsk_krylov_basis_bcsr4(nsk_bcsr4_t B, int s, const double *v0, double *V, int64_t ldv, nsk_where where)
{
    if (!B) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = B->ctx;
    const int64_t n = 4 * (int64_t)B->nbrows;
    NSK_REQUIRE(ctx, s >= 0 && v0 && V && ldv >= n, "bad arguments");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return NSK_OK;
    if (where == NSK_DEVICE) {
        NSK_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (ldv & 1) == 0, "V columns must be 16-byte aligned");
        if (v0 != V) NSK_CUDA(ctx, cudaMemcpyAsync(V, v0, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        for (int j = 0; j < s; j++) NSK_TRY(spmm_bcsr4_device(B, 1, V + (size_t)j * ldv, ldv, V + (size_t)(j + 1) * ldv, ldv));
        return NSK_OK;
    }
    const int64_t ld = (n + 1) & ~(int64_t)1;
    void *vv;
    NSK_TRY(nsk_stage(ctx, 0, sizeof(double) * (size_t)ld * (s + 1), &vv));
    double *dV = (double *)vv;
    NSK_CUDA(ctx, cudaMemcpyAsync(dV, v0, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    for (int j = 0; j < s; j++) NSK_TRY(spmm_bcsr4_device(B, 1, dV + (size_t)j * ld, ld, dV + (size_t)(j + 1) * ld, ld));
    NSK_CUDA(ctx, cudaMemcpy2DAsync(V, sizeof(double) * ldv, dV, sizeof(double) * ld, sizeof(double) * n, s + 1, cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NSK_OK;
}


// -----------------------------------------------------------------------------------------------
// Matrix powers on the block format: levels[l] = B^(l+1) x.  Replaces SpM2V_BCSR / _OPT / _FMA / _AVX2 (reference
// mpk/SpM2V.cpp:376-801) for k = 2 and extends them to any k.  Per scalar row the chain is the (block, j) order of
// SpMV_BCSR_FMA at every level, so the result is bit-identical to k block products.
//
// Default: k launches of the block product kernel, vectors device-resident in between (measured fastest for FEM block
// operators: 58 entries per row leave a fused kernel's window nothing to reuse from L2 that the products do not already
// get, profiles/r02_configs.txt).  With option mpk_kernel = 5 (or 4) the call runs ONE fused launch of the level
// pipeline instead, on the scalar expansion of the blocks (explicit zeros kept, entry order = (block, j): same bits).
static int bcsr4_expand(nsk_bcsr4_t B)
{
    nsk_ctx_t ctx = B->ctx;
    const int nb = B->nbrows;
    const int64_t nblk = B->nblocks;
    NSK_REQUIRE(ctx, 16 * nblk < (int64_t)1 << 31, "block operator too large for 32-bit row pointers of its scalar expansion");
    std::vector<int> bcol((size_t)nblk);
    std::vector<double> bval((size_t)nblk * 16);
    if (nblk) {
        NSK_CUDA(ctx, cudaMemcpy(bcol.data(), B->d_indcol, sizeof(int) * (size_t)nblk, cudaMemcpyDeviceToHost));
        NSK_CUDA(ctx, cudaMemcpy(bval.data(), B->d_coef, sizeof(double) * 16 * (size_t)nblk, cudaMemcpyDeviceToHost));
    }
    const int n = 4 * nb;
    std::vector<int> ptrow((size_t)n + 1), indcol((size_t)nblk * 16);
    std::vector<double> coef((size_t)nblk * 16);
    ptrow[0] = 0;
    for (int bi = 0; bi < nb; bi++) {
        const int p = B->h_ptrow[bi], q = B->h_ptrow[bi + 1];
        for (int i = 0; i < 4; i++) {
            size_t o = (size_t)16 * p + (size_t)4 * (q - p) * i;
            for (int ia = p; ia < q; ia++)
                for (int j = 0; j < 4; j++, o++) {
                    indcol[o] = 4 * bcol[ia] + j;
                    coef[o] = bval[(size_t)16 * ia + 4 * i + j];
                }
            ptrow[4 * bi + i + 1] = (int)o;
        }
    }
    return nsk_csr_create(ctx, n, n, 16 * nblk, ptrow.data(), indcol.data(), coef.data(), &B->expanded);
}

NSK_API int nsk_bcsr4_mpk(nsk_bcsr4_t B, int k, const double *x, double *const *levels, nsk_mode mode, nsk_where where)
{
    if (!B) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = B->ctx;
    NSK_REQUIRE(ctx, k >= 1 && k <= NSK_MAX_K && x && levels, "bad arguments");
    for (int l = 0; l < k; l++) NSK_REQUIRE(ctx, levels[l] != nullptr, "a level vector is null");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->opt.mpk_kernel == 5 || ctx->opt.mpk_kernel == 4) {
        if (!B->expanded) NSK_TRY(bcsr4_expand(B));
        return nsk_mpk(B->expanded, k, x, levels, mode, where);
    }
    if (where == NSK_DEVICE) {
        const double *src = x;
        for (int l = 0; l < k; l++) {
            NSK_TRY(nsk_spmv_bcsr4(B, src, levels[l], mode, NSK_DEVICE));
            src = levels[l];
        }
        ctx->last_mpk = 1;
        return NSK_OK;
    }
    // host pointers: x in once, the level vectors out while the next product runs is not worth a second stream here
    // (the block product is a small fraction of the PCIe time): products back to back, then k copies out
    const size_t nbytes = sizeof(double) * 4 * (size_t)B->nbrows;
    void *vx;
    NSK_TRY(nsk_stage(ctx, 0, nbytes, &vx));
    NSK_CUDA(ctx, cudaMemcpyAsync(vx, x, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    const double *src = (const double *)vx;
    for (int l = 0; l < k; l++) {
        void *vl;
        NSK_TRY(nsk_stage(ctx, 2 + (l & 1), nbytes, &vl));
        NSK_TRY(nsk_spmv_bcsr4(B, src, (double *)vl, mode, NSK_DEVICE));
        NSK_CUDA(ctx, cudaMemcpyAsync(levels[l], vl, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
        src = (const double *)vl;
    }
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->last_mpk = 1;
    return NSK_OK;
}
