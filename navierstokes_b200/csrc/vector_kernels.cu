// vector_kernels.cu -- fused dot / axpy / Gram kernels for the Krylov solvers, plus the public
// vector entry points of nsk.h.
//
// Replaces the reference's only dot+axpy code, orthogonalize (mpk/2SpMV.cpp:3-11,
// mpk/SpMVmulti.cpp:146-151), and the parity metric norm2 / rel_error (mpk/utils.cpp:131-143).
// HBM-bound streaming: 128-bit loads, grid = SM-count multiple, reductions are two-stage and
// deterministic (fixed grid -> fixed association; the last CTA to finish folds the per-CTA
// partials in index order), results stay in device scalar slots so solvers never synchronise.
#include <math.h>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"

using namespace nskptx;

constexpr int VEC_THREADS = 256;

static int vec_grid(nsk_ctx_t ctx, int64_t n)
{
    int64_t want = (n + (int64_t)VEC_THREADS * 4 - 1) / ((int64_t)VEC_THREADS * 4);
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// Block-wide deterministic sum of NS values per thread; result valid in thread 0.
template <int NS>
__device__ __forceinline__ void block_reduce(double (&v)[NS], double *sh /* NS * 8 doubles */)
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
            for (int w = 0; w < VEC_THREADS / 32; w++) t += sh[s * 8 + w];
            v[s] = t;
        }
    }
}

// Publishes NS per-CTA partials and lets the last CTA fold them into out[0..NS).
template <int NS>
__device__ __forceinline__ void grid_finish(double (&v)[NS], double *partials, unsigned int *ticket, double *out)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) partials[(size_t)s * NSK_MAX_PARTIALS + blockIdx.x] = v[s];
        __threadfence();
        unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double t = 0.0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += 32)
                t += ld_cg_f64(partials + (size_t)s * NSK_MAX_PARTIALS + b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) out[s] = t;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// MODE 0: sum a*b ; MODE 1: sum (a-b)^2
template <int MODE>
__global__ void __launch_bounds__(VEC_THREADS) dot_kernel(int64_t n, const double *__restrict__ a,
                                                          const double *__restrict__ b, double *partials,
                                                          unsigned int *ticket, double *out)
{
    __shared__ double sh[8];
    double acc[1] = {0.0};
    // 128-bit loads need 16-byte aligned operands; otherwise everything goes through the scalar tail
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const int64_t n2 = vec ? (n >> 1) : 0;
    const double2 *a2 = reinterpret_cast<const double2 *>(a);
    const double2 *b2 = reinterpret_cast<const double2 *>(b);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 u = a2[i], w = b2[i];
        if (MODE == 0) {
            acc[0] = __fma_rn(u.x, w.x, acc[0]);
            acc[0] = __fma_rn(u.y, w.y, acc[0]);
        } else {
            double d0 = u.x - w.x, d1 = u.y - w.y;
            acc[0] = __fma_rn(d0, d0, acc[0]);
            acc[0] = __fma_rn(d1, d1, acc[0]);
        }
    }
    for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u = a[i], w = b[i];
        if (MODE == 0) acc[0] = __fma_rn(u, w, acc[0]);
        else { double d = u - w; acc[0] = __fma_rn(d, d, acc[0]); }
    }
    block_reduce<1>(acc, sh);
    grid_finish<1>(acc, partials, ticket, out);
}

__global__ void __launch_bounds__(VEC_THREADS) axpy_kernel(int64_t n, double alpha, const double *d_alpha,
                                                           double scale, const double *__restrict__ x,
                                                           double *__restrict__ y)
{
    const double a = d_alpha ? scale * (*d_alpha) : alpha;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int64_t n2 = vec ? (n >> 1) : 0;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    double2 *y2 = reinterpret_cast<double2 *>(y);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 u = x2[i], w = y2[i];
        w.x = __fma_rn(a, u.x, w.x);
        w.y = __fma_rn(a, u.y, w.y);
        y2[i] = w;
    }
    for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = __fma_rn(a, x[i], y[i]);
}

// Gram block: G[i][j] = <V_i, V_j>, i <= j, for m <= 9 vectors -> up to 45 sums per pass.
struct GramPtrs {
    const double *v[12];
};
template <int M>
__global__ void __launch_bounds__(VEC_THREADS) gram_kernel(int64_t n, GramPtrs P, double *partials,
                                                           unsigned int *ticket, double *out)
{
    constexpr int NS = M * (M + 1) / 2;
    __shared__ double sh[NS * 8];
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) acc[s] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v[M];
#pragma unroll
        for (int a = 0; a < M; a++) v[a] = P.v[a][i];
        int s = 0;
#pragma unroll
        for (int a = 0; a < M; a++)
#pragma unroll
            for (int b = a; b < M; b++) { acc[s] = __fma_rn(v[a], v[b], acc[s]); s++; }
    }
    block_reduce<NS>(acc, sh);
    grid_finish<NS>(acc, partials, ticket, out);
}

// Wide Gram pass for the s-step block (M = 7 or 9 vectors, 28 / 45 sums): the accumulators alone take 2 NS registers per
// thread, so the register file -- not the grid -- decides how many bytes an SM has in flight.  gram_kernel<9> (one element
// per thread and trip, 128 registers, 512 threads per SM -> 36 KB in flight per SM) reaches 0.59 of the HBM copy peak
// (profiles/r02_cg_launches.txt); here a thread keeps 2 x E2 elements of every vector in flight (E2 128-bit loads per
// vector, all issued before the first fma) and one CTA of GW_THREADS threads owns the SM's registers.
// Needs 16-byte aligned vectors; the tail (n odd) is one scalar element.  Summation order differs from gram_kernel (as it
// does between any two grids): the Gram block is compared with tolerances, like every tree-summed reduction here.
constexpr int GW_THREADS = 384;
template <int M, int E2>
__global__ void __launch_bounds__(GW_THREADS, 1) gram_wide_kernel(int64_t n, GramPtrs P, double *partials, unsigned int *ticket,
                                                                  double *out)
{
    constexpr int NS = M * (M + 1) / 2;
    constexpr int NW = GW_THREADS / 32;
    __shared__ double sh[NS * NW];
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) acc[s] = 0.0;
    const int64_t n2 = n >> 1;  // double2 elements
    const int64_t chunk = (int64_t)GW_THREADS * E2;
    const int64_t stride = (int64_t)gridDim.x * chunk;
    for (int64_t base = (int64_t)blockIdx.x * chunk + threadIdx.x; base < n2; base += stride) {
        double2 v[E2][M];
#pragma unroll
        for (int e = 0; e < E2; e++) {
            const int64_t i = base + (int64_t)e * GW_THREADS;
#pragma unroll
            for (int a = 0; a < M; a++)
                v[e][a] = i < n2 ? reinterpret_cast<const double2 *>(P.v[a])[i] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int e = 0; e < E2; e++) {
            int s = 0;
#pragma unroll
            for (int a = 0; a < M; a++)
#pragma unroll
                for (int b = a; b < M; b++) {
                    acc[s] = __fma_rn(v[e][a].x, v[e][b].x, acc[s]);
                    acc[s] = __fma_rn(v[e][a].y, v[e][b].y, acc[s]);
                    s++;
                }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        int s = 0;
#pragma unroll
        for (int a = 0; a < M; a++)
#pragma unroll
            for (int b = a; b < M; b++) { acc[s] = __fma_rn(P.v[a][n - 1], P.v[b][n - 1], acc[s]); s++; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; s++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
        if (lane == 0) sh[s * NW + warp] = acc[s];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double t = 0.0;
            for (int w = 0; w < NW; w++) t += sh[s * NW + w];
            acc[s] = t;
        }
    }
    grid_finish<NS>(acc, partials, ticket, out);
}

// ---- launchers --------------------------------------------------------------------------------
int nsk_launch_dot(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, int slot)
{
    dot_kernel<0><<<vec_grid(ctx, n), VEC_THREADS, 0, ctx->stream>>>(n, a, b, ctx->d_partials, ctx->d_ticket,
                                                                      ctx->d_scalars + slot);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

int nsk_launch_diff_norm2sq(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, int slot)
{
    dot_kernel<1><<<vec_grid(ctx, n), VEC_THREADS, 0, ctx->stream>>>(n, a, b, ctx->d_partials, ctx->d_ticket,
                                                                      ctx->d_scalars + slot);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

int nsk_launch_axpy(nsk_ctx_t ctx, int64_t n, double alpha, const double *x, double *y)
{
    axpy_kernel<<<vec_grid(ctx, n), VEC_THREADS, 0, ctx->stream>>>(n, alpha, nullptr, 1.0, x, y);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

int nsk_launch_axpy_dev(nsk_ctx_t ctx, int64_t n, const double *d_alpha, double scale, const double *x, double *y)
{
    axpy_kernel<<<vec_grid(ctx, n), VEC_THREADS, 0, ctx->stream>>>(n, 0.0, d_alpha, scale, x, y);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

// Writes the upper triangle (row-major packed: (0,0),(0,1)..(0,m-1),(1,1)..) to d_scalars[slot0..].
int nsk_launch_gram(nsk_ctx_t ctx, int64_t n, int m, const double *const *vptrs, int slot0)
{
    NSK_REQUIRE(ctx, m >= 1 && m <= 9, "gram supports 1..9 vectors per call");
    NSK_REQUIRE(ctx, slot0 >= 0 && slot0 + m * (m + 1) / 2 <= NSK_NSCALARS, "scalar slots exhausted");
    GramPtrs P;
    for (int i = 0; i < 12; i++) P.v[i] = i < m ? vptrs[i] : nullptr;
    int grid = vec_grid(ctx, n * 4);
    double *out = ctx->d_scalars + slot0;
    // s-step blocks on a long vector: the wide kernel (one CTA per SM, 4 elements of every vector in flight per thread)
    bool aligned = true;
    for (int i = 0; i < m; i++) aligned = aligned && (reinterpret_cast<uintptr_t>(vptrs[i]) & 15) == 0;
    if ((m == 9 || m == 7) && aligned && n >= (int64_t)1 << 16 && ctx->opt.gram_wide >= 0) {
        const int g = ctx->prop.multiProcessorCount;
        if (m == 9) gram_wide_kernel<9, 2><<<g, GW_THREADS, 0, ctx->stream>>>(n, P, ctx->d_partials, ctx->d_ticket, out);
        else gram_wide_kernel<7, 2><<<g, GW_THREADS, 0, ctx->stream>>>(n, P, ctx->d_partials, ctx->d_ticket, out);
        ctx->launches++;
        NSK_CUDA(ctx, cudaGetLastError());
        return NSK_OK;
    }
#define GRAM_CASE(M) case M: gram_kernel<M><<<grid, VEC_THREADS, 0, ctx->stream>>>(n, P, ctx->d_partials, ctx->d_ticket, out); break;
    switch (m) {
        GRAM_CASE(1) GRAM_CASE(2) GRAM_CASE(3) GRAM_CASE(4) GRAM_CASE(5)
        GRAM_CASE(6) GRAM_CASE(7) GRAM_CASE(8) GRAM_CASE(9)
    }
#undef GRAM_CASE
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

int nsk_read_scalars(nsk_ctx_t ctx, int slot0, int count, double *out)
{
    NSK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + slot0, ctx->d_scalars + slot0, sizeof(double) * count,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) out[i] = ctx->h_scalars[slot0 + i];
    return NSK_OK;
}

// ---- public entry points ------------------------------------------------------------------------
namespace {
// stages host vectors on the device when where == NSK_HOST
struct Staged {
    nsk_ctx_t ctx;
    int status = NSK_OK;
    const double *in(int slot, const double *p, int64_t n, nsk_where where)
    {
        if (where == NSK_DEVICE || status != NSK_OK) return p;
        void *d = nullptr;
        status = nsk_stage(ctx, slot, sizeof(double) * (size_t)n, &d);
        if (status != NSK_OK) return nullptr;
        if (cudaMemcpyAsync(d, p, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            nsk_set_error(ctx, "H2D copy failed");
            status = NSK_ERR_CUDA;
        }
        return (const double *)d;
    }
};
const int SLOT_PUBLIC = 0;  // scalar slots 0..7 belong to the public one-shot calls
// The public reductions sum over the communicator's ranks (x, y are the owned parts of distributed vectors) unless the
// caller asked for rank-local results (option local_reductions: a check on one rank only must not enter a collective).
inline int vec_allreduce(nsk_ctx_t ctx, int slot0, int count)
{
    if (ctx->opt.local_reductions) return NSK_OK;
    return nsk_comm_allreduce_slots(ctx, slot0, count);
}
}  // namespace

NSK_API int nsk_dot(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, double *result, nsk_where where)
{
    if (!ctx || !result) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, n >= 0 && (n == 0 || (a && b)), "bad vector");
    Staged S{ctx};
    const double *da = S.in(0, a, n, where), *db = (b == a) ? da : S.in(1, b, n, where);
    NSK_TRY(S.status);
    NSK_TRY(nsk_launch_dot(ctx, n, da, db, SLOT_PUBLIC));
    NSK_TRY(vec_allreduce(ctx, SLOT_PUBLIC, 1));
    return nsk_read_scalars(ctx, SLOT_PUBLIC, 1, result);
}

NSK_API int nsk_norm2(nsk_ctx_t ctx, int64_t n, const double *x, double *result, nsk_where where)
{
    double s = 0.0;
    NSK_TRY(nsk_dot(ctx, n, x, x, &s, where));
    *result = sqrt(s);
    return NSK_OK;
}

NSK_API int nsk_rel_error(nsk_ctx_t ctx, int64_t n, const double *ref, const double *test, double *result,
                          nsk_where where)
{
    if (!ctx || !result) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, n >= 0 && (n == 0 || (ref && test)), "bad vector");
    Staged S{ctx};
    const double *da = S.in(0, ref, n, where), *db = S.in(1, test, n, where);
    NSK_TRY(S.status);
    NSK_TRY(nsk_launch_diff_norm2sq(ctx, n, da, db, SLOT_PUBLIC));
    NSK_TRY(nsk_launch_dot(ctx, n, da, da, SLOT_PUBLIC + 1));
    NSK_TRY(vec_allreduce(ctx, SLOT_PUBLIC, 2));
    double v[2];
    NSK_TRY(nsk_read_scalars(ctx, SLOT_PUBLIC, 2, v));
    *result = sqrt(v[0]) / sqrt(v[1]);
    return NSK_OK;
}

NSK_API int nsk_axpy(nsk_ctx_t ctx, int64_t n, double a, const double *x, double *y, nsk_where where)
{
    if (!ctx) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, n >= 0 && (n == 0 || (x && y)), "bad vector");
    if (where == NSK_DEVICE) return nsk_launch_axpy(ctx, n, a, x, y);
    Staged S{ctx};
    const double *dx = S.in(0, x, n, where);
    double *dy = const_cast<double *>(S.in(1, y, n, where));
    NSK_TRY(S.status);
    NSK_TRY(nsk_launch_axpy(ctx, n, a, dx, dy));
    NSK_CUDA(ctx, cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NSK_OK;
}

NSK_API int nsk_orthogonalize(nsk_ctx_t ctx, int64_t n, const double *x, double *y, double alpha, double *beta,
                              nsk_where where)
{
    if (!ctx) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, n >= 0 && (n == 0 || (x && y)), "bad vector");
    Staged S{ctx};
    const double *dx = S.in(0, x, n, where);
    double *dy = const_cast<double *>(S.in(1, y, n, where));
    NSK_TRY(S.status);
    NSK_TRY(nsk_launch_dot(ctx, n, dx, dy, SLOT_PUBLIC));
    NSK_TRY(vec_allreduce(ctx, SLOT_PUBLIC, 1));
    // y += (-alpha * beta) * x with beta read on the device: no host round trip between dot and axpy
    NSK_TRY(nsk_launch_axpy_dev(ctx, n, ctx->d_scalars + SLOT_PUBLIC, -alpha, dx, dy));
    if (where == NSK_HOST)
        NSK_CUDA(ctx, cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (beta || where == NSK_HOST) {
        double b = 0.0;
        NSK_TRY(nsk_read_scalars(ctx, SLOT_PUBLIC, 1, &b));
        if (beta) *beta = b;
    }
    return NSK_OK;
}

// orthonormalize_against_basis (reference mpk/2SpMV.cpp:13-28): modified Gram-Schmidt sweep of y against every basis
// vector in order -- each projection uses the already updated y -- then ||y||_2, which the reference computes and
// drops (it never scales y); here it is returned.  Two launches per basis vector, the dot stays on the device.
NSK_API int nsk_orthonormalize_against_basis(nsk_ctx_t ctx, int64_t n, int m, const double *const *basis, double *y,
                                             double *norm, nsk_where where)
{
    if (!ctx) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, n >= 0 && m >= 0 && (m == 0 || basis) && (n == 0 || y), "bad arguments");
    const size_t nb = sizeof(double) * (size_t)n;
    double *dy = y;
    double *dbuf = nullptr;
    if (where == NSK_HOST) {
        void *vy = nullptr, *vb = nullptr;
        NSK_TRY(nsk_stage(ctx, 1, nb, &vy));
        NSK_TRY(nsk_stage(ctx, 0, nb, &vb));
        dy = (double *)vy;
        dbuf = (double *)vb;
        NSK_CUDA(ctx, cudaMemcpyAsync(dy, y, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int j = 0; j < m; j++) {
        const double *dx = basis[j];
        if (where == NSK_HOST) {
            NSK_CUDA(ctx, cudaMemcpyAsync(dbuf, basis[j], nb, cudaMemcpyHostToDevice, ctx->stream));
            dx = dbuf;
        }
        NSK_TRY(nsk_launch_dot(ctx, n, dy, dx, SLOT_PUBLIC));
        NSK_TRY(vec_allreduce(ctx, SLOT_PUBLIC, 1));
        NSK_TRY(nsk_launch_axpy_dev(ctx, n, ctx->d_scalars + SLOT_PUBLIC, -1.0, dx, dy));  // y -= <y,x> x
    }
    NSK_TRY(nsk_launch_dot(ctx, n, dy, dy, SLOT_PUBLIC));
    NSK_TRY(vec_allreduce(ctx, SLOT_PUBLIC, 1));
    if (where == NSK_HOST) NSK_CUDA(ctx, cudaMemcpyAsync(y, dy, nb, cudaMemcpyDeviceToHost, ctx->stream));
    double yy = 0.0;
    NSK_TRY(nsk_read_scalars(ctx, SLOT_PUBLIC, 1, &yy));  // also completes the copy above
    if (norm) *norm = sqrt(yy);
    return NSK_OK;
}

NSK_API int nsk_gram(nsk_ctx_t ctx, int64_t n, int m, const double *const *V, double *G, nsk_where where)
{
    if (!ctx || !V || !G) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, m >= 1 && m <= 9, "gram supports 1..9 vectors");
    const double *dv[12];
    if (where == NSK_HOST) {
        void *d = nullptr;
        NSK_TRY(nsk_stage(ctx, 0, sizeof(double) * (size_t)n * (size_t)m, &d));
        for (int i = 0; i < m; i++) {
            double *di = (double *)d + (size_t)i * (size_t)n;
            NSK_CUDA(ctx, cudaMemcpyAsync(di, V[i], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
            dv[i] = di;
        }
    } else {
        for (int i = 0; i < m; i++) dv[i] = V[i];
    }
    const int ns = m * (m + 1) / 2;
    const int slot0 = 64;
    NSK_TRY(nsk_launch_gram(ctx, n, m, dv, slot0));
    NSK_TRY(vec_allreduce(ctx, slot0, ns));
    double tri[45];
    NSK_TRY(nsk_read_scalars(ctx, slot0, ns, tri));
    int s = 0;
    for (int a = 0; a < m; a++)
        for (int b = a; b < m; b++) {
            G[a * m + b] = tri[s];
            G[b * m + a] = tri[s];
            s++;
        }
    return NSK_OK;
}
