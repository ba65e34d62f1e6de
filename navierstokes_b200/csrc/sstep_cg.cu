// sstep_cg.cu -- s-step (communication-avoiding) conjugate gradients on top of the matrix-powers kernel.
//
// The reference has no CG at all (its only "s-step" code is a monomial-basis builder,
// src/kernels/spmm_avx2.c:112-168, and a solver stub, src/sstepgmres.c:126-149), so parity is against the
// restatement kept with the tests (oracle.scg, numpy) -- "parity unpinned" as far as the reference is concerned.
//
// Formulation (Carson/Demmel CA-CG, monomial basis).  One OUTER step advances s CG iterations:
//   1. V = [p, A p, ..., A^s p | r, A r, ..., A^(s-1) r]          ONE two-vector matrix-powers call (depth s for both):
//                                                                 every tile of the operator is streamed once for p
//                                                                 and r together; a distributed slab exchanges one
//                                                                 depth-s halo per vector
//   2. G = V^T V   ((2s+1)^2, symmetric: 45 sums for s = 4)       one pass over the 2s+1 vectors (gram_kernel),
//                                                                 ONE all-reduce per s iterations
//   3. s inner iterations on (2s+1)-long coordinate vectors       one thread: alpha_j = r'Gr' / p'G(Bp'),
//      x' += alpha p';  r' -= alpha B p';  p' = r' + beta p'       B = the shift that maps V c to A V c
//   4. x += V x';  r = V r';  p = V p'                            one pass: 2s+1 reads + x, three writes
// The stopping test ||r||/||b|| <= tol uses the Gram-recurrence residual inside step 3, so the iteration count is
// exact (the block is cut at the converged inner iteration); the caller verifies the true residual.
#include <math.h>

#include <vector>

#include "nsk_internal.h"

int nsk_mpk_device2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                    double *const *d_levels2, nsk_mode mode);  // mpk.cu

constexpr int SCG_MAXS = 4;               // 2s+1 <= 9 vectors per Gram pass
constexpr int SCG_MAXM = 2 * SCG_MAXS + 1;
// scalar slots (ctx->d_scalars): Gram upper triangle, coordinates, status
enum {
    SG_G = 32,               // 45
    SG_XC = 80,              // x' (9)
    SG_RC = 90,              // r' (9)
    SG_PC = 100,             // p' (9)
    SG_RR = 110,             // current <r, r> (recurrence)
    SG_BB = 111,             // <b, b>
    SG_DONE = 112,           // 1 once converged
    SG_ITERS = 113,          // CG iterations performed so far
    SG_BREAK = 114           // 1 if a denominator was not positive (loss of definiteness / basis breakdown)
};

// Step 3: single thread.  G arrives as the packed upper triangle written by gram_kernel.
__global__ void scg_inner_kernel(double *scal, int s, double tol2, int max_iters)
{
    const int m = 2 * s + 1;
    double G[SCG_MAXM][SCG_MAXM];
    int q = 0;
    for (int a = 0; a < m; a++)
        for (int b = a; b < m; b++) { G[a][b] = scal[SG_G + q]; G[b][a] = G[a][b]; q++; }
    double xc[SCG_MAXM], rc[SCG_MAXM], pc[SCG_MAXM], w[SCG_MAXM], t[SCG_MAXM];
    for (int i = 0; i < m; i++) { xc[i] = 0.0; rc[i] = 0.0; pc[i] = 0.0; }
    pc[0] = 1.0;
    rc[s + 1] = 1.0;
    double rr = G[s + 1][s + 1];
    int iters = (int)scal[SG_ITERS];
    // done and brk are sticky: an outer step launched after convergence (or a breakdown) -- the host reads the status only
    // every other outer step -- leaves x' = 0, r' = e_r, p' = e_p, i.e. the update is the identity
    double done = scal[SG_DONE], brk = scal[SG_BREAK];
    const double bb = scal[SG_BB];
    if (done != 0.0 || brk != 0.0) {  // latched: identity coordinates, status and <r, r> stay as they were
        for (int i = 0; i < m; i++) { scal[SG_XC + i] = xc[i]; scal[SG_RC + i] = rc[i]; scal[SG_PC + i] = pc[i]; }
        return;
    }
    for (int j = 0; j < s && done == 0.0 && brk == 0.0 && iters < max_iters; j++) {
        // w = B p': shift inside the P block (0..s) and inside the R block (s+1..2s)
        for (int i = 0; i < m; i++) w[i] = 0.0;
        for (int i = 0; i < s; i++) w[i + 1] = pc[i];
        for (int i = s + 1; i < 2 * s; i++) w[i + 1] = pc[i];
        double denom = 0.0;
        for (int a = 0; a < m; a++) {
            double ga = 0.0;
            for (int b = 0; b < m; b++) ga += G[a][b] * w[b];
            denom += pc[a] * ga;
        }
        if (!(denom > 0.0) || !(rr > 0.0)) { brk = 1.0; break; }
        const double alpha = rr / denom;
        for (int i = 0; i < m; i++) { xc[i] += alpha * pc[i]; rc[i] -= alpha * w[i]; }
        double rr_new = 0.0;
        for (int a = 0; a < m; a++) {
            t[a] = 0.0;
            for (int b = 0; b < m; b++) t[a] += G[a][b] * rc[b];
            rr_new += rc[a] * t[a];
        }
        iters++;
        if (rr_new <= tol2 * bb) {
            rr = rr_new < 0.0 ? 0.0 : rr_new;
            done = 1.0;
            break;
        }
        const double beta = rr_new / rr;
        for (int i = 0; i < m; i++) pc[i] = rc[i] + beta * pc[i];
        rr = rr_new;
    }
    for (int i = 0; i < m; i++) { scal[SG_XC + i] = xc[i]; scal[SG_RC + i] = rc[i]; scal[SG_PC + i] = pc[i]; }
    scal[SG_RR] = rr;
    scal[SG_DONE] = done;
    scal[SG_ITERS] = (double)iters;
    scal[SG_BREAK] = brk;
}

struct ScgPtrs {
    const double *v[SCG_MAXM];
};

// Step 4.  V_0 = p and V_(s+1) = r are also outputs: every element is read in full before it is written.
template <int M>
__global__ void __launch_bounds__(256) scg_update_kernel(int64_t n, ScgPtrs V, double *__restrict__ x, double *r, double *p,
                                                         const double *__restrict__ scal)
{
    double xc[M], rc[M], pc[M];
#pragma unroll
    for (int i = 0; i < M; i++) { xc[i] = scal[SG_XC + i]; rc[i] = scal[SG_RC + i]; pc[i] = scal[SG_PC + i]; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v[M];
#pragma unroll
        for (int a = 0; a < M; a++) v[a] = V.v[a][i];
        double xs = x[i], rs = 0.0, ps = 0.0;
#pragma unroll
        for (int a = 0; a < M; a++) {
            xs = __fma_rn(xc[a], v[a], xs);
            rs = __fma_rn(rc[a], v[a], rs);
            ps = __fma_rn(pc[a], v[a], ps);
        }
        x[i] = xs;
        r[i] = rs;
        p[i] = ps;
    }
}

// Step 4, wide form: the 3 M coordinates live in shared memory (read back as broadcast LDS operands) instead of 6 M
// registers, and a thread updates two consecutive elements with 128-bit accesses -- 4 x the bytes in flight per SM of the
// kernel above (98 registers, 80 bytes per thread: 0.61 of the HBM copy peak, profiles/r02_cg_launches.txt).  Per element
// the three fma chains run in the same order, so the results are bit-identical to scg_update_kernel.
template <int M>
__global__ void __launch_bounds__(256) scg_update_wide_kernel(int64_t n, ScgPtrs V, double *__restrict__ x, double *r, double *p,
                                                              const double *__restrict__ scal)
{
    __shared__ double c[3][M];
    if (threadIdx.x < M) {
        c[0][threadIdx.x] = scal[SG_XC + threadIdx.x];
        c[1][threadIdx.x] = scal[SG_RC + threadIdx.x];
        c[2][threadIdx.x] = scal[SG_PC + threadIdx.x];
    }
    __syncthreads();
    const volatile double *cv = &c[0][0];
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v[M];
#pragma unroll
        for (int a = 0; a < M; a++) v[a] = reinterpret_cast<const double2 *>(V.v[a])[i];
        double2 xs = reinterpret_cast<const double2 *>(x)[i], rs = make_double2(0.0, 0.0), ps = make_double2(0.0, 0.0);
        asm volatile("" ::: "memory");  // all M + 1 loads are issued before the first coordinate is read
#pragma unroll
        for (int a = 0; a < M; a++) {
            // volatile: re-read per trip (broadcast LDS) -- hoisted out of the loop the 27 coordinates cost 54 registers
            const double xa = cv[a], ra = cv[M + a], pa = cv[2 * M + a];
            xs.x = __fma_rn(xa, v[a].x, xs.x);
            xs.y = __fma_rn(xa, v[a].y, xs.y);
            rs.x = __fma_rn(ra, v[a].x, rs.x);
            rs.y = __fma_rn(ra, v[a].y, rs.y);
            ps.x = __fma_rn(pa, v[a].x, ps.x);
            ps.y = __fma_rn(pa, v[a].y, ps.y);
        }
        reinterpret_cast<double2 *>(x)[i] = xs;
        reinterpret_cast<double2 *>(r)[i] = rs;
        reinterpret_cast<double2 *>(p)[i] = ps;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        double v[M];
#pragma unroll
        for (int a = 0; a < M; a++) v[a] = V.v[a][i];
        double xs = x[i], rs = 0.0, ps = 0.0;
#pragma unroll
        for (int a = 0; a < M; a++) {
            xs = __fma_rn(c[0][a], v[a], xs);
            rs = __fma_rn(c[1][a], v[a], rs);
            ps = __fma_rn(c[2][a], v[a], ps);
        }
        x[i] = xs;
        r[i] = rs;
        p[i] = ps;
    }
}

int nsk_scg_device(nsk_csr_t A, const double *d_b, double *d_x, double tol, int maxit, int s, int *iters,
                   double *relres)
{
    nsk_ctx_t ctx = A->ctx;
    NSK_REQUIRE(ctx, s >= 2 && s <= SCG_MAXS, "s-step CG supports s = 2..4 (2s+1 <= 9 vectors per Gram pass)");
    const int n = nsk_csr_owned_rows(A);
    const size_t nloc = (size_t)A->n_cols;  // local vector length (owned + ghost rings for a distributed slab)
    const size_t vec_bytes = ((sizeof(double) * nloc + 255) / 256) * 256;
    const int m = 2 * s + 1;
    double *scal = ctx->d_scalars;

    // workspace: p, r (local vectors), s levels of p, s levels of r (the two-vector powers kernel runs both to depth s;
    // A^s r is not part of the basis and simply not used)
    unsigned char *ws = nullptr;
    {
        void *vws = nullptr;  // grow-only staging slot of the context: repeated solves do not re-allocate 1.2 GB
        if (nsk_stage(ctx, 6, vec_bytes * (size_t)(2 * s + 2), &vws) != NSK_OK) {
            nsk_set_error(ctx, "s-step CG workspace (%d vectors of %zu bytes) does not fit", 2 * s + 2, vec_bytes);
            return NSK_ERR_ALLOC;
        }
        ws = reinterpret_cast<unsigned char *>(vws);
    }
    auto vec = [&](int i) { return reinterpret_cast<double *>(ws + vec_bytes * (size_t)i); };
    double *p = vec(0), *r = vec(1);
    std::vector<double *> lvP(s), lvR(s);
    for (int l = 0; l < s; l++) lvP[l] = vec(2 + l);
    for (int l = 0; l < s; l++) lvR[l] = vec(2 + s + l);
    int status = NSK_OK;
    auto fail = [&](int st) { cudaStreamSynchronize(ctx->stream); return st; };
#define SCG_TRY(call) do { status = (call); if (status != NSK_OK) return fail(status); } while (0)
#define SCG_CUDA(call) do { if ((call) != cudaSuccess) { nsk_set_error(ctx, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); return fail(NSK_ERR_CUDA); } } while (0)

    const size_t nb = sizeof(double) * (size_t)n;
    SCG_CUDA(cudaMemsetAsync(ws, 0, vec_bytes * (size_t)(2 * s + 2), ctx->stream));
    SCG_CUDA(cudaMemsetAsync(d_x, 0, nb, ctx->stream));
    SCG_CUDA(cudaMemcpyAsync(r, d_b, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    SCG_CUDA(cudaMemcpyAsync(p, d_b, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    SCG_CUDA(cudaMemsetAsync(scal + SG_G, 0, sizeof(double) * (SG_BREAK + 1 - SG_G), ctx->stream));
    SCG_TRY(nsk_launch_dot(ctx, n, d_b, d_b, SG_BB));
    SCG_TRY(nsk_comm_allreduce_slots(ctx, SG_BB, 1));
    double h[8];
    SCG_TRY(nsk_read_scalars(ctx, SG_RR, 5, h));
    const double bb = h[SG_BB - SG_RR];
    if (bb == 0.0) {
        if (iters) *iters = 0;
        if (relres) *relres = 0.0;
        return fail(NSK_OK);
    }

    ScgPtrs V;
    const double *gram_ptrs[SCG_MAXM];
    for (int i = 0; i < SCG_MAXM; i++) V.v[i] = nullptr;
    V.v[0] = p;
    for (int l = 0; l < s; l++) V.v[1 + l] = lvP[l];
    V.v[s + 1] = r;
    for (int l = 0; l + 1 < s; l++) V.v[s + 2 + l] = lvR[l];
    for (int i = 0; i < m; i++) gram_ptrs[i] = V.v[i];

    const double tol2 = tol * tol;
    int64_t want = ((int64_t)n + 511) / 512;
    const int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    // 128-bit form of the block update: every vector 16-byte aligned (the workspace is; x is the caller's)
    bool wide = ctx->opt.scg_update_wide >= 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
    for (int i = 0; i < m; i++) wide = wide && (reinterpret_cast<uintptr_t>(V.v[i]) & 15) == 0;
    int done_iters = 0;
    bool converged = false, broke = false;
    double rr = bb;
    // The status (iterations, converged, breakdown) is latched on the device; the host reads it every SCG_POLL outer
    // steps (and whenever the iteration cap could have been reached), so the stream is drained once per 2s iterations.
    constexpr int SCG_POLL = 2;
    int since = 0;
    while (done_iters < maxit && !converged && !broke) {
        SCG_TRY(nsk_mpk_device2(A, s, p, lvP.data(), r, lvR.data(), NSK_EXACT_FMA));
        SCG_TRY(nsk_launch_gram(ctx, n, m, gram_ptrs, SG_G));
        SCG_TRY(nsk_comm_allreduce_slots(ctx, SG_G, m * (m + 1) / 2));
        scg_inner_kernel<<<1, 1, 0, ctx->stream>>>(scal, s, tol2, maxit);
        if (wide) {
            switch (m) {
                case 5: scg_update_wide_kernel<5><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
                case 7: scg_update_wide_kernel<7><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
                default: scg_update_wide_kernel<9><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
            }
        } else {
            switch (m) {
                case 5: scg_update_kernel<5><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
                case 7: scg_update_kernel<7><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
                default: scg_update_kernel<9><<<grid, 256, 0, ctx->stream>>>(n, V, d_x, r, p, scal); break;
            }
        }
        ctx->launches += 2;
        SCG_CUDA(cudaGetLastError());
        if (++since < SCG_POLL && done_iters + (since + 1) * s <= maxit) continue;
        since = 0;
        SCG_TRY(nsk_read_scalars(ctx, SG_RR, 5, h));  // one small D2H + sync per SCG_POLL * s iterations
        rr = h[0];
        converged = h[SG_DONE - SG_RR] != 0.0;
        done_iters = (int)h[SG_ITERS - SG_RR];
        broke = h[SG_BREAK - SG_RR] != 0.0;
    }
#undef SCG_TRY
#undef SCG_CUDA
    if (iters) *iters = done_iters;
    if (relres) *relres = sqrt((rr < 0.0 ? 0.0 : rr) / bb);
    cudaStreamSynchronize(ctx->stream);
    if (broke && !converged) {
        nsk_set_error(ctx, "s-step CG: basis breakdown (non-positive curvature in the Gram recurrence) after %d iterations",
                      done_iters);
        return NSK_ERR_NOT_CONVERGED;
    }
    return converged ? NSK_OK : NSK_ERR_NOT_CONVERGED;
}
