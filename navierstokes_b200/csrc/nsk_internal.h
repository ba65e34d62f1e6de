// nsk_internal.h -- shared declarations of the library's translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/nsk.h"

#define NSK_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing ------------------------------------------------------------------------
void nsk_set_error(nsk_ctx_t ctx, const char *fmt, ...);

#define NSK_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            nsk_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(e__));                                           \
            return NSK_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define NSK_REQUIRE(ctx, cond, msg)                                                           \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            nsk_set_error((ctx), "%s:%d invalid argument: %s", __FILE__, __LINE__, (msg));    \
            return NSK_ERR_INVALID;                                                           \
        }                                                                                     \
    } while (0)

#define NSK_TRY(call)                                                                         \
    do {                                                                                      \
        int s__ = (call);                                                                     \
        if (s__ != NSK_OK) return s__;                                                        \
    } while (0)

// ---- context -------------------------------------------------------------------------------
struct nsk_comm_s;  // comm.cpp

struct nsk_options {
    int64_t spmv_kernel = 0;      // 0 auto (packed when applicable, else stream), 1 scalar (thread/row from global),
                                  // 2 stream (CSR slices by TMA, x gathered from global), 3 packed (x runs staged too),
                                  // 4 sliced-ELL tiles (sell.cu)
    int64_t packed_variant = 0;   // 0 default, else 1 + index into the packed kernel table
    int64_t spmv_ctas_per_sm = 0; // 0 = kernel default
    int64_t mpk_kernel = 0;       // 0 auto (5 for operators made of pattern tiles, else 4 if the operator packs, else 1),
                                  // 1 = k separate products, 4 = level pipeline on the packed format, 5 = on sliced-ELL tiles
    int64_t host_overlap = 1;     // host-pointer powers calls: copy level l out while level l+1.. are computed
    int64_t pk_flags = 1;         // packed kernel switches (see PkParams::flags); default 1: evict-first / streaming hints
    int64_t pk_timing = 0;        // 1: packed kernel records its stage cycle and prints per-level averages (debug)
    int64_t pipe_w0_pct = 0;      // share weight of level 0's team relative to 100 for every other level; 0 = default
    int64_t pipe_bp_global = -1;  // back-pressure of the packed level pipeline: 0 = level l held by l+1, 1 = level 0 held
                                  // by k-1 (one window for the whole pipeline), < 0 = default (1)
    int64_t pipe_interleave = 1;  // 1: level = blockIdx % k (each SM hosts every level), 0: level = blockIdx / team
    int64_t stream_exact_kind = 0; // exact modes of the stream kernel: 0 auto, 1 thread-per-row gathers, 2 gather to smem first
    int64_t stream_variant = 0;   // 0 auto, else 1 + index into the stream kernel table
    int64_t wave_slack_pct = -1;  // extra level skew, % of (resident CTAs x stages / k) tiles; <0 = default 150
    int64_t wave_l2_pct = 0;      // share of L2 the wavefront window may occupy, %; 0 = default 80
    int64_t scg_update_wide = 0;  // < 0: s-step block update through the scalar kernel (comparison)
    int64_t mpk_auto_explicit = 0; // automatic strategy on operators stored as explicit-column tiles (unstructured FEM): fused from
                                   // 1 M rows (0), from this many rows (> 0), never (< 0)
    int64_t gram_wide = 0;        // < 0: s-step Gram blocks always through the one-element-per-thread kernel (comparison)
    int64_t local_reductions = 0; // 1: nsk_dot / norm2 / rel_error / orthogonalize / gram do not all-reduce over the communicator
    int64_t halo_push = 1;        // distributed operators: registered vectors exchange their halo by pushing over NVLink
                                  // peer memory (0: always NCCL)
    int64_t bcsr_batch = 0;       // block product: blocks whose loads are issued together per thread (0 = default 2; 1, 2, 4)
    // sliced-ELL kernel (sell.cu)
    int64_t sell_chunk = 0;       // consecutive tiles a CTA takes per item; 0 = default (2 fused, 4 single product)
    int64_t sell_geom = 0;        // 2 = operators stored with one global pattern still take the masked consumer path (A/B)
    int64_t sell_ctas_per_sm = 0; // 0 = what the occupancy calculator allows
    int64_t sell_max_ctas = 0;    // > 0: cap of the persistent grid (tests: many items per CTA)
    int64_t sell_flags = -1;      // < 0 default (3): bit 0 eviction / streaming hints, bit 1 L2 prefetch of level 0's tiles
    int64_t sell_pf_dist = 0;     // items ahead the L2 prefetch runs; 0 = default 2
    int64_t sell_rows = 0;        // streaming kernel: rows of a tile per consumer thread (0 = default 1, 2)
    int64_t sell_tma = 0;         // all-pattern operators: coefficient stages per CTA of the staged kernel (0 = default 4,
                                  // < 0 = use the register kernels)
    int64_t sell_stream = 0;      // all-pattern operators: tiles in flight per consumer thread (0 = default 3, < 0 = the
                                  // item-at-a-time kernel)
};

struct nsk_ipc_mapping {  // a peer allocation mapped into this process (dist.cu)
    unsigned char handle[64];
    void *base;
};

struct nsk_ctx_s {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host copies of host-pointer calls, overlapped with later products
    cudaEvent_t copy_event[NSK_MAX_K] = {};  // (created on first use)
    cudaDeviceProp prop{};
    uint64_t launches = 0;
    int last_spmv = 0;  // kernel family of the last product: 1 scalar, 2 stream (CSR), 3 packed, 4 sliced-ELL tiles
    int last_sell[8] = {};  // plan of the last sliced-ELL launch (queries sell_uniform_width, sell_reach, sell_lead, ...)
    int last_mpk = 0;   // strategy of the last powers call: 1 levels, 4 packed level pipeline, 5 sliced-ELL level pipeline
    std::string last_error;
    nsk_options opt;
    // scratch for reductions: partial sums + ticket + result slots (device) and pinned mirror
    double *d_partials = nullptr;  // NSK_MAX_PARTIALS * NSK_RED_SLOTS
    unsigned int *d_ticket = nullptr;
    double *d_scalars = nullptr;   // NSK_NSCALARS device scalars (dot results, CG state)
    double *h_scalars = nullptr;   // pinned mirror
    // L2 scrub buffer
    void *d_flush = nullptr;
    size_t flush_bytes = 0;
    // staging buffers for NSK_HOST calls (grown on demand)
    void *d_stage[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    nsk_comm_s *comm = nullptr;
    std::vector<nsk_ipc_mapping> ipc_maps;
};
void nsk_ipc_close_all(nsk_ctx_t ctx);  // dist.cu

constexpr int NSK_MAX_PARTIALS = 4096;  // max blocks of a reducing kernel
constexpr int NSK_RED_SLOTS = 96;       // max simultaneous sums per reducing kernel (Gram 9x9 = 81)
constexpr int NSK_NSCALARS = 256;

int nsk_stage(nsk_ctx_t ctx, int slot, size_t bytes, void **p);  // grow-only device staging buffer

// ---- CSR operator --------------------------------------------------------------------------
// One unit of work of the streaming kernels: a run of consecutive rows whose col/val slice fits
// one shared-memory stage.
struct nsk_tile {
    int row0;   // first row
    int nrows;  // rows in the tile (>= 1)
    int nz0;    // ptrow[row0]
    int nz1;    // ptrow[row0 + nrows]
};

struct nsk_tiling {
    int tile_nnz = 0;   // capacity of a stage in nonzeros
    int tile_rows = 0;  // max rows per tile
    int ntiles = 0;
    int nlong = 0;      // tiles consisting of one row longer than tile_nnz
    nsk_tile *d_tiles = nullptr;
    std::vector<nsk_tile> h_tiles;
};

struct nsk_dist_s;  // dist.cu

struct nsk_csr_s {
    nsk_ctx_t ctx = nullptr;
    int n = 0, n_cols = 0;
    int64_t nnz = 0;
    int *d_ptrow = nullptr;    // n+1 (+pad)
    int *d_indcol = nullptr;   // nnz (+pad)
    double *d_coef = nullptr;  // nnz (+pad)
    double mean_row = 0.0;
    int max_row = 0;
    std::vector<nsk_tiling *> tilings;  // one per tile geometry in use (built on demand, kept)
    std::vector<int> breaks;            // local rows at which a tile must end (segment starts of a distributed slab)
    std::vector<int> row_rank;          // distributed: position of each local row in GLOBAL row order (else empty)
    // MPK scratch: k level vectors when the caller passes host memory, wavefront flags, ...
    int *d_flags = nullptr;
    size_t flags_count = 0;
    nsk_dist_s *dist = nullptr;
};

struct nsk_bcsr4_s {
    nsk_ctx_t ctx = nullptr;
    int nbrows = 0;
    int64_t nblocks = 0;
    int *d_ptrow = nullptr;
    int *d_indcol = nullptr;
    double *d_coef = nullptr;
    std::vector<int> h_ptrow;         // block row pointers (host copy: sizes of the scalar expansion)
    nsk_csr_t expanded = nullptr;     // scalar CSR with the blocks' explicit zeros, built on the first FUSED powers call
};

// ---- kernel launchers (spmv_kernels.cu) -----------------------------------------------------
// Row range [row_begin,row_end) lets MPK levels and distributed slabs evaluate a prefix of rows.
struct nsk_spmv_args {
    const double *x = nullptr;
    double *y = nullptr;
    int row_begin = 0, row_end = 0;
    nsk_mode mode = NSK_EXACT_FMA;
    // optional fused dot: partial sums of w[i]*y[i] over the rows computed (CG: w = x = p)
    const double *dot_w = nullptr;
    int dot_slot = -1;  // index into ctx->d_scalars receiving the finished sum
};
int nsk_launch_spmv(nsk_csr_t A, const nsk_spmv_args &a);
// packed.cu: k = 1 is the plain product (optionally with the fused dot); NSK_ERR_UNSUPPORTED = use the CSR kernels
int nsk_packed_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode, const int *level_rows,
                   const double *dot_w, int dot_slot);
int nsk_packed_run2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                    double *const *d_levels2, nsk_mode mode, const int *level_rows);
bool nsk_packed_applicable(nsk_csr_t A);
// sell.cu: sliced-ELL tiles, any CSR operator; k = 1 is the plain product; NSK_ERR_UNSUPPORTED = use another path
int nsk_sell_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode, const int *level_rows,
                 const double *dot_w, int dot_slot);
int nsk_sell_run2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                  double *const *d_levels2, nsk_mode mode, const int *level_rows);
bool nsk_sell_applicable(nsk_csr_t A);
bool nsk_sell_uniform(nsk_csr_t A);
bool nsk_sell_explicit_staged(nsk_csr_t A);  // explicit columns, lengths + columns of every tile fit a stage  // all tiles pattern tiles of one width (stencils, regular bands)
size_t nsk_sell_bytes(nsk_csr_t A);
void nsk_sell_free(nsk_csr_t A);
int nsk_sell_check_error(nsk_csr_t A);
size_t nsk_packed_bytes(nsk_csr_t A);
void nsk_packed_free(nsk_csr_t A);
int nsk_get_tiling(nsk_csr_t A, int t_nnz, int t_rows, const nsk_tiling **out);  // cached per geometry
void nsk_free_tilings(nsk_csr_t A);
void nsk_stream_kernel_config(nsk_ctx_t ctx, double mean_row, int *tile_nnz, int *tile_rows);

// ---- vector kernels (vector_kernels.cu) ------------------------------------------------------
// All results land in ctx->d_scalars[slot]; nothing synchronises.
int nsk_launch_dot(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, int slot);
int nsk_launch_diff_norm2sq(nsk_ctx_t ctx, int64_t n, const double *a, const double *b, int slot);
int nsk_launch_axpy(nsk_ctx_t ctx, int64_t n, double alpha, const double *x, double *y);
int nsk_launch_axpy_dev(nsk_ctx_t ctx, int64_t n, const double *d_alpha, double scale,
                        const double *x, double *y);
int nsk_launch_gram(nsk_ctx_t ctx, int64_t n, int m, const double *const *d_vptrs, int slot0);
int nsk_read_scalars(nsk_ctx_t ctx, int slot0, int count, double *out);  // syncs the stream

// ---- comm (comm.cpp) -------------------------------------------------------------------------
int nsk_comm_allreduce_slots(nsk_ctx_t ctx, int slot0, int count);  // no-op without a communicator
bool nsk_comm_active(nsk_ctx_t ctx);
int nsk_comm_sendrecv(nsk_ctx_t ctx, int npeers, const int *peer, const double *const *sendbuf,
                      const int *sendcount, double *const *recvbuf, const int *recvcount);
