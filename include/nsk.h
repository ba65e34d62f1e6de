/*
 * nsk.h -- C ABI of the B200-native SpMV / matrix-powers / Krylov kernels.
 *
 * This is the drop-in boundary for the hot path of aantoine890/navierstokes (tree mpk/): the
 * entry points below are what a foreign-function binding (ctypes, cgo, JNI, a PETSc
 * MATOP_MULT callback, or the C++ shims in include/nsk_spmv_compat.hpp) binds instead of the
 * reference's CPU kernels.  Plain C types only: opaque handles, raw pointers, sizes, int status.
 *
 *   reference interface (file:line under /root/reference)        replaced by
 *   -----------------------------------------------------------  --------------------------------
 *   struct csrmatrix {n,nnz,ptrow,indcol,coef}  mpk/SpMV.h:18-24  nsk_csr_create / nsk_csr_destroy
 *   SpMV_CSR, _OPT, _FMA, _AVX2                 mpk/SpMV.h:55-58  nsk_spmv (mode selects arithmetic)
 *   SpM2V_CSR*, SpM2V (k=2)     mpk/SpM2V.cpp:80,137,195,279      nsk_mpk  (k = 2)
 *   SpM3V, SpM4V (k=3,4)        mpk/SpMVmulti0.cpp:132,191        nsk_mpk  (k = 3, 4; any k <= NSK_MAX_K)
 *   Generate{1st,2nd,3rd}layer  mpk/SpM2V.cpp:5, SpMVmulti0:106,157  plan built inside nsk_csr_create
 *   struct bcsr4x4_matrix, SpMV_BCSR*           mpk/SpMV.h:26-33,61-64  nsk_bcsr4_create / nsk_spmv_bcsr4
 *   norm2, rel_error                            mpk/utils.cpp:131-143  nsk_norm2 / nsk_rel_error
 *   orthogonalize (dot + axpy)                  mpk/2SpMV.cpp:3-11     nsk_orthogonalize, nsk_dot, nsk_axpy
 *   orthonormalize_against_basis                mpk/2SpMV.cpp:13-28    nsk_orthonormalize_against_basis
 *   flush_cache                                 mpk/utils.cpp:146-154  nsk_flush_l2
 *   COO2CSR / generate_CSR, generate_BCSR4      mpk/utils.cpp:5-127    nsk_coo2csr, nsk_coo2bcsr4 (host)
 *   the .mtx reader inlined in every driver     mpk/SpM2V.cpp:815-852  nsk_mtx_read (host)
 *   BuildKrylovBasis_AVX2 / MatMatMult_SeqBAIJ_4_AVX2 (s-step basis, several right-hand sides)
 *                                               src/kernels/spmm_avx2.c:7-168  nsk_mpk_multi
 *   (no CG in the reference; north star asks for it)             nsk_cg (classical and s-step)
 *
 * Conventions
 *   - One nsk_ctx per process and per GPU (one process per GPU; ranks are joined by nsk_comm_init).
 *   - Every function returns NSK_OK (0) or a negative nsk_status; nsk_last_error() gives detail.
 *     The reference's kernels are `void` and never fail; the C++ shims abort loudly on an error
 *     because there is NO CPU fallback.
 *   - `where` says whether vector pointers are host or device memory.  With NSK_HOST the call
 *     copies inputs to the GPU, runs, copies results back and returns after they have landed (the
 *     reference's synchronous contract).  With NSK_DEVICE the call only enqueues work on the
 *     context's stream; use nsk_ctx_sync() or your own stream/event to wait.
 *   - Matrix arrays passed to *_create are host pointers (like the reference's std::vector data);
 *     they are copied to HBM once and never referenced afterwards.
 *   - Outputs are fully overwritten (the reference zeroes then accumulates, mpk/SpMV.cpp:13);
 *     inputs are never written.
 */
#ifndef NSK_H
#define NSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSK_VERSION 100          /* 0.1.0 */
#define NSK_MAX_K 16             /* largest matrix-powers depth per call */
#define NSK_UNIQUE_ID_BYTES 128  /* == sizeof(ncclUniqueId) */

typedef enum {
    NSK_OK = 0,
    NSK_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, unsorted ptrow ...) */
    NSK_ERR_CUDA = -2,         /* a CUDA runtime call or kernel failed */
    NSK_ERR_NO_DEVICE = -3,    /* no usable sm_100 device: there is no CPU fallback */
    NSK_ERR_ALLOC = -4,        /* host or device allocation failed */
    NSK_ERR_COMM = -5,         /* NCCL missing or a collective failed */
    NSK_ERR_UNSUPPORTED = -6,  /* valid request that this build does not implement */
    NSK_ERR_NOT_CONVERGED = -7 /* nsk_cg hit maxit (x still holds the last iterate) */
} nsk_status;

/* Arithmetic of one row of y = A x.
 * NSK_EXACT_FMA    y_i = fma(a_ip, x_p, ...fma(a_i1, x_1, fma(a_i0, x_0, +0.0))) in CSR storage
 *                  order -- bit-identical to SpMV_CSR_FMA / SpMV_CSR_OPT (mpk/SpMV.cpp:23-56).
 * NSK_EXACT_MULADD same order, product and sum rounded separately -- bit-identical to the code
 *                  g++ emits for the inner reduction of SpM2V_CSR_OPT on the development host and
 *                  to SSE2 evaluation of SpMV_CSR (mpk/SpMV.cpp:6-20).
 * NSK_FAST         any association (sub-warp partial sums + shuffle tree); within 1e-12 relative
 *                  (reference metric rel_error, mpk/utils.cpp:138-143) of NSK_EXACT_FMA. */
typedef enum { NSK_EXACT_FMA = 0, NSK_EXACT_MULADD = 1, NSK_FAST = 2 } nsk_mode;

typedef enum { NSK_HOST = 0, NSK_DEVICE = 1 } nsk_where;

typedef struct nsk_ctx_s *nsk_ctx_t;
typedef struct nsk_csr_s *nsk_csr_t;
typedef struct nsk_bcsr4_s *nsk_bcsr4_t;

/* ---- library / context ---------------------------------------------------------------------- */
int nsk_version(void);
const char *nsk_strerror(int status);
/* Text of the last failure on this context (or the last context-less failure when ctx == NULL). */
const char *nsk_last_error(nsk_ctx_t ctx);

/* Binds the calling process to CUDA device `device` and creates the work stream.
 * Fails with NSK_ERR_NO_DEVICE when there is no GPU or it is not compute capability 10.x. */
int nsk_ctx_create(int device, nsk_ctx_t *ctx);
int nsk_ctx_destroy(nsk_ctx_t ctx);
/* Use an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL restores ours. */
int nsk_ctx_set_stream(nsk_ctx_t ctx, void *cuda_stream);
void *nsk_ctx_get_stream(nsk_ctx_t ctx);
int nsk_ctx_sync(nsk_ctx_t ctx);
/* Number of kernels this library has launched on this context since creation. */
uint64_t nsk_ctx_launch_count(nsk_ctx_t ctx);
/* Device facts: sm_count, l2_bytes, smem_per_block_optin, total HBM bytes. */
int nsk_ctx_device_info(nsk_ctx_t ctx, int *sm_count, int64_t *l2_bytes, int *smem_optin,
                        int64_t *hbm_bytes);
/* Tuning knobs (name -> integer).  Unknown names give NSK_ERR_INVALID.  Defaults are what bench.py runs.
 *   spmv_kernel      0 auto (sliced-ELL tiles for operators made of pattern tiles, packed when the operator packs, else
 *                    stream) | 1 scalar | 2 stream (CSR) | 3 packed | 4 sliced-ELL tiles
 *   mpk_kernel       0 auto (fused level pipeline: sliced-ELL tiles for operators made of pattern tiles, the packed
 *                    format for other operators that pack, else k launches) | 1 k launches | 4 level pipeline
 *                    (packed) | 5 level pipeline (sliced-ELL tiles: any operator whose rows are not too ragged)
 *   sell_tma         all-pattern operators: coefficient stages per CTA of the staged kernel (0 = default 3: four CTAs per
 *                    SM; 4: three CTAs; < 0 = the register kernels); explicit-column operators: entries of a row loaded
 *                    per round trip (0 = default 16 with two CTAs per SM; 8: 8 with three CTAs; 17: 16 with three CTAs)
 *                    | sell_chunk tiles per item (0 = default) |
 *                    sell_stream, sell_rows (register kernels) | sell_ctas_per_sm | sell_flags | sell_pf_dist
 *   packed_variant, stream_variant   0 default, n = table entry n-1 of that kernel
 *   wave_l2_pct      share of L2 the fused kernels' window may occupy (0 = default: 92 sliced-ELL pattern operators, 88 otherwise, 70 packed)
 *   wave_slack_pct   explicit window slack (< 0 = size it from the L2 budget)
 *   pipe_bp_global, pipe_interleave, pipe_w0_pct, pk_flags, stream_exact_kind, spmv_ctas_per_sm
 *                    experiment switches documented next to nsk_options in csrc/nsk_internal.h
 *   pk_flags         bit 0 (default on) cache hints for data nobody re-reads | bit 1 poll without sleeping | bit 2
 *                    publish with red.release
 *   pk_timing        1: the fused kernels print their stage-cycle / dependency-wait breakdown to stderr (debugging aid)
 *   mpk_auto_explicit  automatic strategy on unstructured operators (explicit-column tiles): fused from 1 M rows when
 *                    min(k, 4) levels fit one launch (0 = default), from `value` rows (> 0), never (< 0)
 *   local_reductions 1: nsk_dot / nsk_norm2 / nsk_rel_error / nsk_orthogonalize / nsk_gram return rank-local results although
 *                    a communicator is attached (no collective)
 *   bcsr_batch       block product: blocks whose loads are issued together (0 = default 2; 1, 2, 4)
 *   gram_wide, scg_update_wide   < 0: the one-element-per-thread forms of the s-step Gram pass / block update (comparison)
 *   halo_push        0: distributed operators always exchange halos through NCCL (default 1: registered vectors push) */
int nsk_ctx_set_option(nsk_ctx_t ctx, const char *name, int64_t value);

/* CUDA-event timing on the context's stream (what bench.py brackets its timed regions with). */
/* Introspection for tests / benchmarks: "last_spmv_kernel" (1 scalar, 2 stream over CSR, 3 packed, 4 sliced-ELL tiles),
 * "last_mpk_strategy" (1 k launches, 4 level pipeline over the packed format, 5 level pipeline over sliced-ELL tiles),
 * "launches". */
int nsk_ctx_query(nsk_ctx_t ctx, const char *name, int64_t *value);
int nsk_event_create(nsk_ctx_t ctx, void **event);
int nsk_event_destroy(nsk_ctx_t ctx, void *event);
int nsk_event_record(nsk_ctx_t ctx, void *event);
/* Waits for `stop`, then returns milliseconds between the two events. */
int nsk_event_elapsed_ms(nsk_ctx_t ctx, void *start, void *stop, float *ms);

/* ---- memory --------------------------------------------------------------------------------- */
int nsk_malloc(nsk_ctx_t ctx, size_t bytes, void **dptr);
int nsk_free(nsk_ctx_t ctx, void *dptr);
int nsk_host_alloc(nsk_ctx_t ctx, size_t bytes, void **hptr); /* pinned */
int nsk_host_free(nsk_ctx_t ctx, void *hptr);
/* kind: 0 = host->device, 1 = device->host, 2 = device->device; asynchronous on the ctx stream. */
int nsk_memcpy(nsk_ctx_t ctx, void *dst, const void *src, size_t bytes, int kind);
int nsk_memset0(nsk_ctx_t ctx, void *dptr, size_t bytes);
/* GPU analogue of flush_cache (mpk/utils.cpp:146-154): overwrites a scratch buffer larger than L2. */
int nsk_flush_l2(nsk_ctx_t ctx);

/* ---- ingest: the step in front of the hot path (host only, no GPU) ------------------------------ */
/* COO -> CSR with the semantics of COO2CSR / generate_CSR (reference mpk/utils.cpp:97-127, 5-43): columns
 * ascending inside a row, a repeated (i,j) is dropped (first wins).  ptrow[nrow+1], indcol/coef sized nnz.
 * Returns the number of entries kept, or a negative nsk_status. */
int64_t nsk_coo2csr(int nrow, int64_t nnz, const int *irow, const int *jcol, const double *val, int *ptrow,
                    int *indcol, double *coef);
/* COO -> 4x4 block CSR with the semantics of generate_BCSR4 (reference mpk/utils.cpp:45-95): nrow/4 block rows,
 * block columns in first-appearance order, row-major blocks with explicit zeros, a repeated (i,j) overwrites.
 * Call with indcol == NULL to get the block count, then with ptrow[nrow/4+1], indcol[nblk], coef[16*nblk]. */
int64_t nsk_coo2bcsr4(int nrow, int64_t nnz, const int *irow, const int *jcol, const double *val, int *ptrow,
                      int *indcol, double *coef);
/* Matrix Market coordinate reader with the quirks of the reference's drivers (mpk/SpM2V.cpp:815-852): banner
 * line skipped, '%' lines skipped, "rows cols nnz", entries 1-based, values parsed as FLOAT then widened.
 * Arrays are malloc'ed; release with nsk_mtx_free. */
int nsk_mtx_read(const char *path, int *nrow, int64_t *nnz, int **irow, int **jcol, double **val);
void nsk_mtx_free(int *irow, int *jcol, double *val);

/* ---- ordering (host; the reference has none; north star: "RCM-ordered blocks") ----------------- */
/* Reverse Cuthill-McKee on the pattern of A + A^T: perm[new] = old.  Level-set BFS from a pseudo-peripheral node of
 * every connected component, neighbours by ascending degree, the order reversed. */
int nsk_rcm(int n, const int *ptrow, const int *indcol, int *perm);
/* B = P A P^T (row new = row perm[new] of A, columns renumbered and sorted ascending, values with their entries).
 * Output arrays sized n+1 / nnz by the caller; coef / coef_out may be NULL (pattern only). */
int nsk_csr_permute(int n, const int *ptrow, const int *indcol, const double *coef, const int *perm, int *ptrow_out,
                    int *indcol_out, double *coef_out);
/* max |i - j| over the entries of A. */
int64_t nsk_csr_bandwidth(int n, const int *ptrow, const int *indcol);

/* ---- the packed format's host half (no GPU): used by the CPU test-suite to check the packer --------------- */
/* Packs a CSR operator for table entry `variant` of the packed kernel (0-based); nsk_pack_host_why returns "" or
 * the reason it does not pack; nsk_pack_host_expand rebuilds CSR from the blobs (global columns through the tiles'
 * x runs) so a test can compare it entry for entry with the input. */
int nsk_pack_host_create(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                         int variant, void **handle);
const char *nsk_pack_host_why(void *handle);
int64_t nsk_pack_host_bytes(void *handle);
int nsk_pack_host_expand(void *handle, int *ptrow, int *indcol, double *coef, int *max_runs, int *max_xlen);
/* CPU model of the fused kernel's protocol on the schedule the GPU path would build (forward dependencies, window
 * back-pressure, per-group completion counters, in-order CTAs with `stages` open items, seeded random interleaving).
 * The model also checks data readiness: when an item opens, every tile its nonzeros really read must be complete at
 * the level below.  Returns the number of items that never became runnable (0 = the schedule is sound), a negative
 * nsk_status, or -1000000 - v when v reads would have seen rows not yet produced. */
long long nsk_pack_host_simulate(void *handle, int k, int lead_slack_tiles, int resident, int w0_pct, int bp_global,
                                 int interleave, int stages, const int *level_rows, unsigned seed,
                                 long long *items_out, int *reach_out, int ghi_bias /* 0; < 0 weakens the forward
                                 dependencies by that many groups, for negative tests */);
void nsk_pack_host_destroy(void *handle);

/* ---- the sliced-ELL format's host half (no GPU; csrc/sell.cu) ------------------------------------------------ */
/* Cuts a CSR operator into tiles of 256 consecutive rows stored slot-major per 32-row slice.  A tile whose rows all follow
 * one column pattern (slot e of row r references column r + rel[e]; rows may lack slots) stores the pattern once and a
 * slot mask per row -- 8 bytes per nonzero, no per-entry index; every other tile keeps explicit 32-bit columns.
 * nsk_sell_host_why returns "" or the reason the operator is refused (rows too ragged for slot-major slices);
 * nsk_sell_host_expand rebuilds CSR from the tiles so a test can compare it entry for entry with the input. */
int nsk_sell_host_create(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                         void **handle);
const char *nsk_sell_host_why(void *handle);
int nsk_sell_host_stats(void *handle, int64_t *bytes, int64_t *ntiles, int64_t *pattern_tiles);
int nsk_sell_host_expand(void *handle, int *ptrow, int *indcol, double *coef);
/* Number of slots of the operator's global column pattern (0: none), its offsets, and how many tiles are stored with it. */
int nsk_sell_host_global_pattern(void *handle, int *rel, int64_t *global_tiles);
/* CPU model of the fused sliced-ELL kernel's protocol on the schedule the GPU path would build: items of `chunk`
 * consecutive tiles, forward dependencies and window back-pressure through per-group completion counters, in-order
 * CTAs with up to `ring` open items that finish in any order but are published in order.  Same return convention as
 * nsk_pack_host_simulate; pmax_bias < 0 weakens the forward dependencies by that many tiles (negative tests). */
long long nsk_sell_host_simulate(void *handle, int k, int chunk, int lead_slack_tiles, int resident, int w0_pct,
                                 int interleave, int ring, const int *level_rows, unsigned seed, long long *items_out,
                                 int *reach_out, int pmax_bias);
void nsk_sell_host_destroy(void *handle);

/* ---- CSR operator --------------------------------------------------------------------------- */
/* Uploads a square CSR operator (0-based, int32 indices, fp64 values; columns need not be sorted --
 * the row-sequential modes accumulate in storage order whatever it is) and builds the launch plan.
 * n rows, n_cols columns (n_cols >= n; n_cols > n is used for row slabs whose columns include
 * ghost entries).  nnz must equal ptrow[n] and be < 2^31 (the reference's `int nnz`). */
int nsk_csr_create(nsk_ctx_t ctx, int n, int n_cols, int64_t nnz, const int *ptrow,
                   const int *indcol, const double *coef, nsk_csr_t *A);
int nsk_csr_destroy(nsk_csr_t A);
int nsk_csr_shape(nsk_csr_t A, int *n, int *n_cols, int64_t *nnz);
/* Algorithmic bytes of one product / of a depth-k powers call (SURVEY.md 8d definitions). */
int64_t nsk_csr_spmv_bytes(nsk_csr_t A);
int64_t nsk_csr_mpk_bytes(nsk_csr_t A, int k);
/* Bytes of the tile-packed copy of the operator (slot-major tiles with 16-bit local column indices, the
 * form the default kernels stream from HBM), built on first call; 0 when the operator does not pack
 * (its tiles reference x in too many runs) and the CSR kernels are used instead. */
int64_t nsk_csr_packed_bytes(nsk_csr_t A);
/* Bytes of the sliced-ELL tile copy (csrc/sell.cu) the default product / fused-powers kernels stream; 0 when the operator
 * is not stored that way.  Builds the tiles if no product has run yet. */
int64_t nsk_csr_tile_bytes(nsk_csr_t A);

/* y = A x.   x: n_cols doubles, y: n doubles. */
int nsk_spmv(nsk_csr_t A, const double *x, double *y, nsk_mode mode, nsk_where where);

/* levels[l] = A^(l+1) x for l = 0 .. k-1 (the reference's y, z, w, v of SpM2V/SpM3V/SpM4V).
 * Every level is an output of n doubles; x has n doubles (square operator).
 * In the exact modes each level is bit-identical to k successive nsk_spmv calls. */
int nsk_mpk(nsk_csr_t A, int k, const double *x, double *const *levels, nsk_mode mode,
            nsk_where where);

/* Matrix powers of several right-hand sides: levels[v*k + l] = A^(l+1) xs[v], v < nvec.  The s-step / block-Krylov
 * basis builder (the reference's BuildKrylovBasis_AVX2 / MatMatMult_SeqBAIJ_4_AVX2, src/kernels/spmm_avx2.c:7-168):
 * device-resident vectors are processed two per fused launch, so the operator is streamed once per PAIR and level. */
int nsk_mpk_multi(nsk_csr_t A, int k, int nvec, const double *const *xs, double *const *levels, nsk_mode mode,
                  nsk_where where);

/* ---- 4x4 block CSR (the reference's bcsr4x4_matrix, row-major blocks) ------------------------ */
int nsk_bcsr4_create(nsk_ctx_t ctx, int nbrows, int64_t nblocks, const int *ptrow,
                     const int *indcol, const double *coef, nsk_bcsr4_t *B);
int nsk_bcsr4_destroy(nsk_bcsr4_t B);
/* y = B x, both 4*nbrows doubles.  Exact modes follow SpMV_BCSR_FMA's (block, j) order. */
int nsk_spmv_bcsr4(nsk_bcsr4_t B, const double *x, double *y, nsk_mode mode, nsk_where where);
/* levels[l] = B^(l+1) x, l < k: replaces SpM2V_BCSR / _OPT / _FMA / _AVX2 (reference mpk/SpM2V.cpp:376-801) for k = 2,
 * any k <= 16.  Bit-identical to k block products in (block, j) order.  k device-resident launches of the block kernel
 * by default; option mpk_kernel = 5 runs one fused launch of the level pipeline on the scalar expansion of the blocks. */
int nsk_bcsr4_mpk(nsk_bcsr4_t B, int k, const double *x, double *const *levels, nsk_mode mode, nsk_where where);
/* Y = B X for s dense columns (column-major; ldx, ldy >= 4 * nbrows; device pointers: columns 16-byte aligned).
 * Replaces MatMatMult_SeqBAIJ_4_AVX2(A, X, Y, s_step) (reference src/kernels/spmm_avx2.c:7-109): columns in groups of
 * four, per block a four-link fma chain that is then ADDED to the row's sum -- the same grouping, bit for bit.  The
 * reference's horizontal reduction (:93-99) adds four equal lanes and so returns 4 * (B X); this entry returns B X
 * (multiplying by 4 is exact, the tests compare 4 * Y with the literal restatement). */
int nsk_spmm_bcsr4(nsk_bcsr4_t B, int s, const double *X, int64_t ldx, double *Y, int64_t ldy, nsk_where where);
/* V(:,0) = v0, V(:,j+1) = B V(:,j) for j < s (V: s + 1 columns, leading dimension ldv): the monomial Krylov basis of
 * BuildKrylovBasis_AVX2 (reference src/kernels/spmm_avx2.c:112-168), one single-column product per vector. */
int nsk_krylov_basis_bcsr4(nsk_bcsr4_t B, int s, const double *v0, double *V, int64_t ldv, nsk_where where);

/* ---- vectors (device-side dot / axpy family used by the Krylov solvers) ---------------------- */
/* result pointers are HOST doubles; the call returns after the value has landed.
 * COLLECTIVE SEMANTICS: once a communicator is attached to the context (nsk_comm_init with nranks > 1) the vectors are
 * taken as the owned parts of distributed vectors and nsk_dot, nsk_norm2, nsk_rel_error, nsk_orthogonalize,
 * nsk_orthonormalize_against_basis and nsk_gram SUM THEIR REDUCTIONS OVER ALL RANKS (in-stream ncclAllReduce): every
 * rank must make the same call, a call on one rank alone blocks.  For a rank-local result (a check on one rank, a
 * local norm) set nsk_ctx_set_option(ctx, "local_reductions", 1) around the call; nsk_cg always reduces globally. */
int nsk_dot(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, double *result,
            nsk_where where);
int nsk_norm2(nsk_ctx_t ctx, int64_t n, const double *x, double *result, nsk_where where);
/* ||ref - test||_2 / ||ref||_2  (mpk/utils.cpp:138-143) */
int nsk_rel_error(nsk_ctx_t ctx, int64_t n, const double *ref, const double *test, double *result,
                  nsk_where where);
/* y += a x */
int nsk_axpy(nsk_ctx_t ctx, int64_t n, double a, const double *x, double *y, nsk_where where);
/* beta = <x,y>; y -= alpha*beta*x  (mpk/2SpMV.cpp:3-11); *beta may be NULL */
int nsk_orthogonalize(nsk_ctx_t ctx, int64_t n, const double *x, double *y, double alpha,
                      double *beta, nsk_where where);
/* orthonormalize_against_basis (reference mpk/2SpMV.cpp:13-28): for each of the m basis vectors in order,
 * y -= <y, b_j> b_j (modified Gram-Schmidt: every projection sees the updated y); *norm receives ||y||_2 afterwards
 * (the reference computes it and does not scale y; neither does this). */
int nsk_orthonormalize_against_basis(nsk_ctx_t ctx, int64_t n, int m, const double *const *basis, double *y,
                                     double *norm, nsk_where where);
/* G[i*m+j] = <V_i, V_j> for m vectors of length n (s-step Gram block); G is m*m HOST doubles. */
int nsk_gram(nsk_ctx_t ctx, int64_t n, int m, const double *const *V, double *G, nsk_where where);

/* ---- conjugate gradients --------------------------------------------------------------------- */
/* Solves A x = b (A symmetric positive definite), x0 = 0, stops when ||r||_2/||b||_2 <= tol.
 * sstep <= 1: classical CG (3 fused kernels per iteration, scalars stay on the device).
 * sstep  > 1: s-step (communication-avoiding) CG, s = 2..4, monomial basis: two matrix-powers calls
 *             (depth s on p, s-1 on r), ONE Gram reduction / all-reduce and ONE fused update per s
 *             iterations.  iters counts CG iterations (the last block is cut at the converged one).
 * With a communicator attached (nsk_comm_init) A is this rank's row slab created by
 * nsk_csr_create_dist and b, x are the owned parts; dots are all-reduced over NCCL. */
int nsk_cg(nsk_csr_t A, const double *b, double *x, double tol, int maxit, int sstep,
           int *iters, double *relres, nsk_where where);

/* ---- multi-GPU (one process per GPU, NCCL over NVLink) --------------------------------------- */
/* Rank 0 calls nsk_comm_unique_id and ships the 128 bytes to the other ranks by any means
 * (torch.distributed broadcast, MPI, a file); then every rank calls nsk_comm_init. */
int nsk_comm_unique_id(void *id128);
int nsk_comm_init(nsk_ctx_t ctx, int nranks, int rank, const void *id128);
int nsk_comm_destroy(nsk_ctx_t ctx);
int nsk_comm_allreduce_sum(nsk_ctx_t ctx, double *dbuf, int count); /* in place, device buffer */

/* ---- distributed operator: row slabs + depth-k ghost rings ------------------------------------
 * Rank r owns the contiguous global rows [row_starts[r], row_starts[r+1]).  For a depth-K plan the
 * rank stores, besides its own rows, the rows of ghost rings 1..K-1 (so that powers can be evaluated
 * redundantly on the shrinking sets N_{K-1} > ... > N_0 = owned, PA1-style) and addresses x on N_K.
 * Local numbering: [owned | ring 1 | ring 2 | ... | ring K], each ring ascending in global id, so
 * that everything one peer sends for one ring lands in ONE contiguous range of the local vector.
 *
 * Planning is host-only (no GPU, no communicator) and driven by the host layer, which supplies the
 * matrix rows of each ring (from a generator, a file, or by fetching them from their owners):
 *
 *     nsk_plan_create(..., depth, &plan)
 *     for (stage = 0; stage < depth; stage++) {              // exactly `depth` rounds: owned rows, ring 1 .. ring depth-1
 *         nsk_plan_frontier(plan, &cnt, &rows);              // cnt may be 0 (a rank whose ring is empty): still call
 *         nsk_plan_add_rows(plan, cnt, ptr, cols, vals);     // add_rows (NULL arrays are fine when cnt == 0)
 *     }
 *     nsk_plan_finalize(plan)                                // fails unless all `depth` rounds were supplied
 *     for every peer p: nsk_plan_requests(plan, p, ...)  -> ship the id list to p (any transport)
 *                       nsk_plan_add_send(plan, p, ...)  <- the list p shipped to us
 *     nsk_csr_create_dist(ctx, plan, &A)
 */
typedef struct nsk_plan_s *nsk_plan_t;

int nsk_plan_create(int nranks, int rank, const int *row_starts /* nranks+1 */, int depth, nsk_plan_t *plan);
int nsk_plan_destroy(nsk_plan_t plan);
/* Global ids (ascending) of the rows whose matrix rows must be supplied next.  First call: the owned rows; then
 * ring 1, ..., ring depth-1.  *count == 0 means "this ring is empty" before the last round (do not stop: supply the
 * empty round) and "all rounds supplied" after it. */
int nsk_plan_frontier(nsk_plan_t plan, int *count, const int **global_rows);
/* Rows of the current frontier, in frontier order: ptr[count+1], global column ids, values. */
int nsk_plan_add_rows(nsk_plan_t plan, int count, const int *ptr, const int *cols_global, const double *vals);
int nsk_plan_finalize(nsk_plan_t plan);
/* level_rows[l], l = 0..depth-1: leading local rows on which power l+1 of a depth-`depth` call is
 * evaluated (level_rows[depth-1] == n_owned).  ring_start[r], r = 0..depth+1: local index where ring r
 * begins (ring 0 = owned; ring_start[depth+1] == n_cols_local). */
int nsk_plan_sizes(nsk_plan_t plan, int *n_owned, int *n_rows_local, int *n_cols_local, int64_t *nnz,
                   int *level_rows, int *ring_start);
int nsk_plan_ghosts(nsk_plan_t plan, const int **global_ids); /* n_cols_local - n_owned ids, local order */
int nsk_plan_local_csr(nsk_plan_t plan, const int **ptrow, const int **indcol_local, const double **coef);
/* What this rank needs from `peer`: global ids in local (ring-major) order and how many per ring
 * (ring_counts[r-1] for ring r = 1..depth). */
int nsk_plan_requests(nsk_plan_t plan, int peer, int *count, const int **global_ids, int *ring_counts);
/* What `peer` needs from this rank (the list peer obtained from ITS nsk_plan_requests(plan, us)). */
int nsk_plan_add_send(nsk_plan_t plan, int peer, int count, const int *global_ids, const int *ring_counts);

/* Uploads the local operator of a finalized plan and sets up the exchange buffers.  The returned
 * operator has n = n_rows_local rows and n_cols = n_cols_local columns; vectors passed with
 * NSK_DEVICE to nsk_spmv / nsk_mpk / nsk_cg on it are LOCAL vectors of n_cols_local doubles whose
 * first n_owned entries are the owned part (ghost entries are scratch, filled by the exchange).
 * With NSK_HOST the pointers are owned parts only (n_owned doubles). */
int nsk_csr_create_dist(nsk_ctx_t ctx, nsk_plan_t plan, nsk_csr_t *A);
int nsk_csr_owned_rows(nsk_csr_t A); /* n_owned for a distributed operator, n otherwise */
/* Refreshes ghost rings 1..depth of a local device vector (pack kernel + grouped ncclSend/ncclRecv). */
int nsk_halo_exchange(nsk_csr_t A, double *xlocal, int depth);

/* ---- halo push over NVLink peer memory (optional; one node, one process per GPU) ---------------------------------
 * A vector registered with nsk_dist_vector_register is mapped into the neighbours' address spaces (CUDA IPC); for such
 * a vector the depth-k halo of nsk_mpk / nsk_spmv is ONE kernel that stores the entries the neighbours need straight
 * into their ghost slots and raises an arrival flag there, plus a one-warp wait for the neighbours' flags -- no send /
 * receive buffers, no unpack, no NCCL call on the path.  Setup, once per operator (every rank, same order):
 *   nsk_dist_push_flags(A, &f);  nsk_ipc_export(ctx, f, h)          -> send h and, per neighbour p, nsk_dist_recv_layout(A, p, ..)
 *   for every neighbour p: nsk_ipc_import(ctx, h_p, &fp);  nsk_dist_push_peer(A, p, my index in p's peer list,
 *                                                                              p's recv layout for me, fp)
 * then per vector: allocate with nsk_malloc, export, import the neighbours' handles, nsk_dist_vector_register.
 * Unregistered vectors (and the option halo_push = 0) keep the NCCL exchange. */
#define NSK_IPC_HANDLE_BYTES 80 /* CUDA IPC handle of the enclosing allocation + the pointer's offset in it */
int nsk_ipc_export(nsk_ctx_t ctx, void *devptr, unsigned char *handle /* NSK_IPC_HANDLE_BYTES */);
int nsk_ipc_import(nsk_ctx_t ctx, const unsigned char *handle, void **peerptr); /* stays mapped until nsk_ctx_destroy */
int nsk_dist_push_flags(nsk_csr_t A, void **flags);
int nsk_dist_recv_layout(nsk_csr_t A, int peer_rank, int *ring_start, int *ring_count);
int nsk_dist_peer_count(nsk_csr_t A);
int nsk_dist_peer_rank(nsk_csr_t A, int index);
int nsk_dist_push_peer(nsk_csr_t A, int peer_rank, int index_at_peer, const int *peer_ring_start, const int *peer_ring_count,
                       void *peer_flags);
int nsk_dist_vector_register(nsk_csr_t A, double *local, double *const *peer_ptrs);

#ifdef __cplusplus
}
#endif
#endif /* NSK_H */
