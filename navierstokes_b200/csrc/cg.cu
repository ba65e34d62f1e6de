// cg.cu -- conjugate gradients on top of the SpMV / matrix-powers kernels (nsk_cg of nsk.h).
//
// The reference has no CG (its Krylov solver is PETSc GMRES+ILU, src/solve_newton.c:
// 1154-1164); BASELINE.json's north star asks for a pressure-Poisson CG built on the SpMV/MPK path,
// so parity here is against the textbook restatement in oracle/nsk_oracle.c (oracle_cg) --
// "parity unpinned" as far as the reference is concerned.
//
// Classical CG, three launches per iteration, no host synchronisation inside an iteration:
//   1. q = A p  fused with  <p, q>            (streaming SpMV kernel, dot in its epilogue)
//   2. alpha = rr/<p,q>;  x += alpha p;  r -= alpha q;  rr' = <r, r>     (one pass over 4 vectors)
//   3. beta = rr'/rr;  p = r + beta p          (and the convergence latch)
// All scalars live in device slots; the stopping test ||r||/||b|| <= tol is evaluated on the
// device and latches the iteration number, after which kernels 2 and 3 become no-ops, so the
// answer and the iteration count do not depend on how often the host looks at the latch.
// With a communicator (multi-GPU) the two dot results are all-reduced in-stream over NCCL and p's
// ghost entries are refreshed by one halo exchange before the product.
#include <math.h>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"

using namespace nskptx;

int nsk_scg_device(nsk_csr_t A, const double *d_b, double *d_x, double tol, int maxit, int s, int *iters,
                   double *relres);  // sstep_cg.cu

// scalar slots used by CG (ctx->d_scalars)
enum { S_RR0 = 16, S_RR1 = 17, S_PQ = 18, S_BB = 19, S_CONV = 20 /* latched iteration, 0 = running */,
       S_RRFIN = 21 /* <r,r> at the latched iteration: with a communicator the frozen iterations that follow in the same
                       batch still all-reduce the rr slots (a collective cannot be made conditional), which scales them */ };

constexpr int CG_THREADS = 256;

template <int NS>
__device__ __forceinline__ void cg_block_reduce(double (&v)[NS], double *sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; s++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[s] += __shfl_xor_sync(0xffffffffu, v[s], o);
        if (lane == 0) sh[s * 8 + warp] = v[s];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double t = 0.0;
            for (int w = 0; w < CG_THREADS / 32; w++) t += sh[s * 8 + w];
            v[s] = t;
        }
    }
}

// x += alpha p ; r -= alpha q ; rr_new = <r,r>
__global__ void __launch_bounds__(CG_THREADS) cg_update_xr_kernel(int64_t n, const double *__restrict__ p,
                                                                  const double *__restrict__ q,
                                                                  double *__restrict__ x, double *__restrict__ r,
                                                                  double *scal, int rr_in, int rr_out,
                                                                  double *partials, unsigned int *ticket)
{
    if (scal[S_CONV] != 0.0) return;  // converged earlier: freeze
    __shared__ double sh[8];
    __shared__ bool is_last;
    const double alpha = scal[rr_in] / scal[S_PQ];
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q) |
                       reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r)) & 15) == 0;
    const int64_t n2 = vec ? (n >> 1) : 0;
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 *q2 = reinterpret_cast<const double2 *>(q);
    double2 *x2 = reinterpret_cast<double2 *>(x);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 pp = p2[i], qq = q2[i], xx = x2[i], rr = r2[i];
        xx.x = __fma_rn(alpha, pp.x, xx.x);
        xx.y = __fma_rn(alpha, pp.y, xx.y);
        rr.x = __fma_rn(-alpha, qq.x, rr.x);
        rr.y = __fma_rn(-alpha, qq.y, rr.y);
        x2[i] = xx;
        r2[i] = rr;
        acc[0] = __fma_rn(rr.x, rr.x, acc[0]);
        acc[0] = __fma_rn(rr.y, rr.y, acc[0]);
    }
    for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double rr = __fma_rn(-alpha, q[i], r[i]);
        x[i] = __fma_rn(alpha, p[i], x[i]);
        r[i] = rr;
        acc[0] = __fma_rn(rr, rr, acc[0]);
    }
    cg_block_reduce<1>(acc, sh);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc[0];
        __threadfence();
        unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double t = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += ld_cg_f64(partials + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            scal[rr_out] = t;
            *ticket = 0u;
        }
    }
}

// beta = rr'/rr ; p = r + beta p ; latch convergence
__global__ void __launch_bounds__(CG_THREADS) cg_update_p_kernel(int64_t n, const double *__restrict__ r,
                                                                 double *__restrict__ p, double *scal, int rr_in,
                                                                 int rr_out, double tol2, double iter_no)
{
    if (scal[S_CONV] != 0.0) return;
    const double rr_new = scal[rr_out];
    const double beta = rr_new / scal[rr_in];
    const bool converged = rr_new <= tol2 * scal[S_BB];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (!converged) {
        const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(r)) & 15) == 0;
        const int64_t n2 = vec ? (n >> 1) : 0;
        const double2 *r2 = reinterpret_cast<const double2 *>(r);
        double2 *p2 = reinterpret_cast<double2 *>(p);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            double2 rr = r2[i], pp = p2[i];
            pp.x = __fma_rn(beta, pp.x, rr.x);
            pp.y = __fma_rn(beta, pp.y, rr.y);
            p2[i] = pp;
        }
        for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            p[i] = __fma_rn(beta, p[i], r[i]);
    }
    // every CTA has read S_CONV before any CTA can write it?  No: the latch is written by a
    // follow-up single-thread kernel (cg_latch_kernel) so that this grid sees one consistent value.
    (void)iter_no;
}

__global__ void cg_latch_kernel(double *scal, int rr_out, double tol2, double iter_no)
{
    if (scal[S_CONV] == 0.0 && scal[rr_out] <= tol2 * scal[S_BB]) {
        scal[S_CONV] = iter_no;
        scal[S_RRFIN] = scal[rr_out];
    }
}

static int cg_grid(nsk_ctx_t ctx, int64_t n)
{
    int64_t want = (n + (int64_t)CG_THREADS * 2 - 1) / ((int64_t)CG_THREADS * 2);
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

int nsk_halo_exchange_dev(nsk_csr_t A, double *xlocal, int depth, bool allow_push);
int nsk_halo_release_dev(nsk_csr_t A, const double *xlocal, int depth);  // dist.cu

static int cg_device(nsk_csr_t A, const double *d_b, double *d_x, double tol, int maxit, int *iters,
                     double *relres)
{
    nsk_ctx_t ctx = A->ctx;
    const int n = nsk_csr_owned_rows(A);
    const size_t nb = sizeof(double) * (size_t)n;
    void *vr, *vp, *vq;
    NSK_TRY(nsk_stage(ctx, 2, nb, &vr));
    NSK_TRY(nsk_stage(ctx, 3, sizeof(double) * (size_t)A->n_cols, &vp));  // p carries ghost entries
    NSK_TRY(nsk_stage(ctx, 4, nb, &vq));
    double *r = (double *)vr, *p = (double *)vp, *q = (double *)vq;
    double *scal = ctx->d_scalars;

    NSK_CUDA(ctx, cudaMemsetAsync(d_x, 0, nb, ctx->stream));
    NSK_CUDA(ctx, cudaMemcpyAsync(r, d_b, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    NSK_CUDA(ctx, cudaMemsetAsync(p, 0, sizeof(double) * (size_t)A->n_cols, ctx->stream));
    NSK_CUDA(ctx, cudaMemcpyAsync(p, d_b, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    NSK_CUDA(ctx, cudaMemsetAsync(scal + S_RR0, 0, sizeof(double) * 6, ctx->stream));
    NSK_TRY(nsk_launch_dot(ctx, n, d_b, d_b, S_BB));
    NSK_TRY(nsk_comm_allreduce_slots(ctx, S_BB, 1));
    NSK_CUDA(ctx, cudaMemcpyAsync(scal + S_RR0, scal + S_BB, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));

    double h[6];
    NSK_TRY(nsk_read_scalars(ctx, S_RR0, 6, h));
    const double bb = h[S_BB - S_RR0];
    if (bb == 0.0) {
        if (iters) *iters = 0;
        if (relres) *relres = 0.0;
        return NSK_OK;
    }
    const double tol2 = tol * tol;
    const int grid = cg_grid(ctx, n);
    const int check_every = 8;
    int it = 0;
    int conv_iter = 0;
    while (it < maxit && conv_iter == 0) {
        int batch = maxit - it < check_every ? maxit - it : check_every;
        for (int b = 0; b < batch; b++, it++) {
            const int rr_in = S_RR0 + (it & 1), rr_out = S_RR0 + ((it + 1) & 1);
            if (A->dist) NSK_TRY(nsk_halo_exchange_dev(A, p, 1, false));
            nsk_spmv_args a;
            a.x = p;
            a.y = q;
            a.row_begin = 0;
            a.row_end = n;
            a.mode = NSK_EXACT_FMA;
            a.dot_w = p;
            a.dot_slot = S_PQ;
            NSK_TRY(nsk_launch_spmv(A, a));
            NSK_TRY(nsk_comm_allreduce_slots(ctx, S_PQ, 1));
            cg_update_xr_kernel<<<grid, CG_THREADS, 0, ctx->stream>>>(n, p, q, d_x, r, scal, rr_in, rr_out,
                                                                       ctx->d_partials, ctx->d_ticket);
            ctx->launches++;
            NSK_TRY(nsk_comm_allreduce_slots(ctx, rr_out, 1));
            cg_update_p_kernel<<<grid, CG_THREADS, 0, ctx->stream>>>(n, r, p, scal, rr_in, rr_out, tol2,
                                                                      (double)(it + 1));
            cg_latch_kernel<<<1, 1, 0, ctx->stream>>>(scal, rr_out, tol2, (double)(it + 1));
            ctx->launches += 2;
        }
        NSK_CUDA(ctx, cudaGetLastError());
        NSK_TRY(nsk_read_scalars(ctx, S_RR0, 6, h));
        conv_iter = (int)h[S_CONV - S_RR0];
    }
    const int done = conv_iter ? conv_iter : it;
    // rr after `done` iterations sits in slot parity done&1
    const double rr = conv_iter ? h[S_RRFIN - S_RR0] : h[done & 1];
    if (iters) *iters = done;
    if (relres) *relres = sqrt(rr / bb);
    return conv_iter ? NSK_OK : (sqrt(rr / bb) <= tol ? NSK_OK : NSK_ERR_NOT_CONVERGED);
}

NSK_API int nsk_cg(nsk_csr_t A, const double *b, double *x, double tol, int maxit, int sstep, int *iters,
                   double *relres, nsk_where where)
{
    if (!A) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    NSK_REQUIRE(ctx, b && x, "b or x is null");
    NSK_REQUIRE(ctx, tol > 0.0 && maxit >= 0, "bad tolerance / maxit");
    NSK_REQUIRE(ctx, sstep <= 4, "s-step depth above 4 is not supported (2s+1 <= 9 vectors per Gram pass)");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nb = sizeof(double) * (size_t)nsk_csr_owned_rows(A);
    const double *db = b;
    double *dx = x;
    if (where == NSK_HOST) {
        void *vb, *vx;
        NSK_TRY(nsk_stage(ctx, 0, nb, &vb));
        NSK_TRY(nsk_stage(ctx, 1, nb, &vx));
        NSK_CUDA(ctx, cudaMemcpyAsync(vb, b, nb, cudaMemcpyHostToDevice, ctx->stream));
        db = (const double *)vb;
        dx = (double *)vx;
    }
    int s = sstep > 1 ? nsk_scg_device(A, db, dx, tol, maxit, sstep, iters, relres)
                      : cg_device(A, db, dx, tol, maxit, iters, relres);
    if (s != NSK_OK && s != NSK_ERR_NOT_CONVERGED) return s;
    if (where == NSK_HOST) {
        NSK_CUDA(ctx, cudaMemcpyAsync(x, dx, nb, cudaMemcpyDeviceToHost, ctx->stream));
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return s;
}
