// spmv_kernels.cu -- fp64 CSR y = A x for sm_100a.
//
// Replaces SpMV_CSR / _OPT / _FMA / _AVX2 (reference mpk/SpMV.cpp:6-85).  HBM-bound integer+fp64
// streaming work: no tensor cores.  Design:
//
//   * The operator is cut at create time into TILES: runs of consecutive rows whose col/val slice
//     (<= T_NNZ nonzeros, <= T_ROWS rows) fits one shared-memory stage.  In CSR such a slice is
//     contiguous, so one tile = three 1-D bulk async copies (TMA, SASS UBLKCP): ptrow slice,
//     indcol slice, coef slice -- 12 B/nnz arrive in shared memory with no register staging, no
//     L1 pollution, and a deep queue of bytes in flight per SM.
//   * Persistent CTAs, warp-specialised: one producer warp (one elected lane) runs a STAGES-deep
//     mbarrier ring of tiles; NCW consumer warps take a landed tile and reduce rows.
//   * Row reduction flavours (KIND):
//       0  one thread per row, nonzeros in storage order            (exact; best for short rows:
//          stencils -- the gather x[col] of 32 consecutive rows is 32 consecutive doubles per
//          diagonal, i.e. fully coalesced, and the fma chain is bit-identical to the reference)
//       1  G lanes per row + shuffle tree                           (fast mode, long rows)
//       2  all threads gather x[col] into shared memory, then one thread per row runs the
//          sequential chain out of shared memory                    (exact mode, long rows)
//     x is read through the read-only path (L1-allocating), everything else bypasses L1.
//   * Optional fused dot product <w, y> over the rows produced (CG: p.Ap), reduced
//     deterministically (fixed grid -> fixed association) and finished by the last CTA.
//
// A plain thread-per-row / lanes-per-row pair reading straight from global memory is kept as the
// simple path ("scalar" kernels): used for cross-checking the pipeline in tests and selectable
// with the option spmv_kernel=1.  Both are GPU kernels; there is no CPU path.
#include <algorithm>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"
#include "stream_common.cuh"

using namespace nskptx;

// -----------------------------------------------------------------------------------------------
// simple kernels (global memory only)
// -----------------------------------------------------------------------------------------------
template <bool MULADD>
__global__ void __launch_bounds__(256) spmv_scalar_kernel(const int *__restrict__ ptrow,
                                                          const int *__restrict__ indcol,
                                                          const double *__restrict__ coef,
                                                          const double *__restrict__ x,
                                                          double *__restrict__ y, int row_begin,
                                                          int row_end)
{
    int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    int p = ptrow[i], q = ptrow[i + 1];
    double acc = 0.0;
    for (int j = p; j < q; j++) acc = row_op<MULADD>(coef[j], __ldg(x + indcol[j]), acc);
    y[i] = acc;
}

template <int G>
__global__ void __launch_bounds__(256) spmv_vector_kernel(const int *__restrict__ ptrow,
                                                          const int *__restrict__ indcol,
                                                          const double *__restrict__ coef,
                                                          const double *__restrict__ x,
                                                          double *__restrict__ y, int row_begin,
                                                          int row_end)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int i = row_begin + t / G;
    int l = t % G;
    bool valid = i < row_end;
    double acc = 0.0;
    if (valid) {
        int p = ptrow[i], q = ptrow[i + 1];
        for (int j = p + l; j < q; j += G) acc = __fma_rn(coef[j], __ldg(x + indcol[j]), acc);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (valid && l == 0) y[i] = acc;
}

// -----------------------------------------------------------------------------------------------
// streaming kernel
// -----------------------------------------------------------------------------------------------
struct StreamParams {
    const nsk_tile *tiles;
    int tile_begin, tile_end;
    const int *ptrow;
    const int *indcol;
    const double *coef;
    const double *x;
    double *y;
    int row_begin, row_end;
    // fused dot
    const double *dot_w;
    double *partials;
    unsigned int *ticket;
    double *dot_out;
};

template <int T_NNZ, int T_ROWS, int STAGES, int KIND>
constexpr int stream_smem_bytes()
{
    // stages + (KIND 2: gathered-x buffer) + barriers + reduction scratch
    return StageGeom<T_NNZ, T_ROWS>::BYTES * STAGES + (KIND == 2 ? T_NNZ * 8 : 0) + 2 * STAGES * 8 + 64 * 8 + 128;
}

// Long row (more nonzeros than a stage holds): read straight from global memory.
template <bool MULADD, bool FAST>
__device__ __forceinline__ void long_row(const StreamParams &P, int row, int lane, int cwarp, double &dot_acc)
{
    if (cwarp != 0) return;
    int p = P.ptrow[row], q = P.ptrow[row + 1];
    double acc = 0.0;
    if (FAST) {
        for (int j = p + lane; j < q; j += 32) acc = __fma_rn(P.coef[j], __ldg(P.x + P.indcol[j]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    } else if (lane == 0) {
        for (int j = p; j < q; j++) acc = row_op<MULADD>(P.coef[j], __ldg(P.x + P.indcol[j]), acc);
    }
    if (lane == 0) {
        P.y[row] = acc;
        if (P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc, dot_acc);
    }
}

template <int T_NNZ, int T_ROWS, int STAGES, int NCW, int MINB, int KIND, int G, bool MULADD>
__global__ void __launch_bounds__((NCW + 1) * 32, MINB) spmv_stream_kernel(const StreamParams P)
{
    using Geo = StageGeom<T_NNZ, T_ROWS>;
    constexpr int NCT = NCW * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage_base = smem;
    double *xg_s = reinterpret_cast<double *>(smem + Geo::BYTES * STAGES);  // KIND 2 only
    unsigned char *tail = smem + Geo::BYTES * STAGES + (KIND == 2 ? T_NNZ * 8 : 0);
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty = full + STAGES;
    double *red = reinterpret_cast<double *>(empty + STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);      // producer's arrive (+ transaction bytes)
            mbar_init(&empty[s], NCW);   // one arrive per consumer warp
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int first = P.tile_begin + blockIdx.x;
    const int step = gridDim.x;

    if (warp == NCW) {
        // ===== producer: one lane keeps STAGES tiles in flight =====
        if (lane == 0) {
            nsk_tile nxt;
            if (first < P.tile_end) nxt = P.tiles[first];
            int it = 0;
            for (int tile = first; tile < P.tile_end; tile += step, ++it) {
                const nsk_tile t = nxt;
                if (tile + step < P.tile_end) nxt = P.tiles[tile + step];  // prefetch descriptor
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);  // fresh barrier: passes at once
                unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
                int *hdr = reinterpret_cast<int *>(st + Geo::HDR_OFF);
                hdr[0] = t.row0; hdr[1] = t.nrows; hdr[2] = t.nz0; hdr[3] = t.nz1;
                const int nnz = t.nz1 - t.nz0;
                if (nnz > T_NNZ) {
                    mbar_arrive(&full[s]);  // long row: nothing staged
                } else {
                    const int a0 = t.nz0 & ~3, v0 = t.nz0 & ~1, p0 = t.row0 & ~3;
                    const uint32_t cb = (uint32_t)(((t.nz1 - a0) + 3) & ~3) * 4u;
                    const uint32_t vb = (uint32_t)(((t.nz1 - v0) + 1) & ~1) * 8u;
                    const uint32_t pb = (uint32_t)(((t.row0 + t.nrows + 1 - p0) + 3) & ~3) * 4u;
                    mbar_arrive_expect_tx(&full[s], cb + vb + pb);
                    bulk_g2s(st + Geo::PTR_OFF, P.ptrow + p0, pb, &full[s]);
                    if (cb) bulk_g2s(st + Geo::COL_OFF, P.indcol + a0, cb, &full[s]);
                    if (vb) bulk_g2s(st + Geo::VAL_OFF, P.coef + v0, vb, &full[s]);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int ctid = tid;  // consumer warps are warps 0..NCW-1
    double dot_acc = 0.0;
    int it = 0;
    for (int tile = first; tile < P.tile_end; tile += step, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
        const int *hdr = reinterpret_cast<const int *>(st + Geo::HDR_OFF);
        const int row0 = hdr[0], nrows = hdr[1], nz0 = hdr[2], nz1 = hdr[3];
        if (nz1 - nz0 > T_NNZ) {
            if (row0 >= P.row_begin && row0 < P.row_end)
                long_row<MULADD, KIND == 1>(P, row0, lane, warp, dot_acc);
        } else {
            const double *val_s = reinterpret_cast<const double *>(st + Geo::VAL_OFF);
            const int *col_s = reinterpret_cast<const int *>(st + Geo::COL_OFF);
            const int *ptr_s = reinterpret_cast<const int *>(st + Geo::PTR_OFF);
            // element with GLOBAL nonzero index j sits at val_s[j - vo], col_s[j - co]; row r at ptr_s[r - po]
            const int vo = nz0 & ~1, co = nz0 & ~3, po = row0 & ~3;
            if (KIND == 0) {
                for (int r = ctid; r < nrows; r += NCT) {
                    const int row = row0 + r;
                    if (row < P.row_begin || row >= P.row_end) continue;
                    const int p = ptr_s[row - po], q = ptr_s[row + 1 - po];
                    const double acc = row_chain<MULADD, true>(val_s, col_s, p, q, vo, co, P.x);
                    P.y[row] = acc;
                    if (P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc, dot_acc);
                }
            } else if (KIND == 1) {
                constexpr int NGRP = NCT / G;
                const int g = ctid / G, l = ctid % G;
                for (int rb = 0; rb < nrows; rb += NGRP) {
                    const int row = row0 + rb + g;
                    const bool valid = (rb + g) < nrows && row >= P.row_begin && row < P.row_end;
                    double acc = 0.0;
                    if (valid) {
                        const int p = ptr_s[row - po], q = ptr_s[row + 1 - po];
                        for (int j = p + l; j < q; j += G) acc = __fma_rn(val_s[j - vo], __ldg(P.x + col_s[j - co]), acc);
                    }
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (valid && l == 0) {
                        P.y[row] = acc;
                        if (P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc, dot_acc);
                    }
                }
            } else {
                // phase A: every consumer thread gathers x for a strided share of the nonzeros
                for (int j = nz0 + ctid; j < nz1; j += NCT) xg_s[j - nz0] = __ldg(P.x + col_s[j - co]);
                named_bar_sync(1, NCT);
                // phase B: sequential chain per row out of shared memory
                for (int r = ctid; r < nrows; r += NCT) {
                    const int row = row0 + r;
                    if (row < P.row_begin || row >= P.row_end) continue;
                    const int p = ptr_s[row - po], q = ptr_s[row + 1 - po];
                    double acc = 0.0;
#pragma unroll 4
                    for (int j = p; j < q; j++) acc = row_op<MULADD>(val_s[j - vo], xg_s[j - nz0], acc);
                    P.y[row] = acc;
                    if (P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc, dot_acc);
                }
                named_bar_sync(1, NCT);  // xg_s is reused by the next tile
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (P.dot_w) {
        // deterministic: lanes -> warp (xor tree), warps in order, CTAs in order (last CTA finishes)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_xor_sync(0xffffffffu, dot_acc, o);
        if (lane == 0) red[warp] = dot_acc;
        named_bar_sync(2, NCT);
        if (warp == 0) {
            __shared__ bool is_last;
            if (lane == 0) {
                double s = 0.0;
                for (int w = 0; w < NCW; w++) s += red[w];
                P.partials[blockIdx.x] = s;
                __threadfence();
                unsigned int done = atomicAdd(P.ticket, 1u);
                is_last = (done == gridDim.x - 1);
            }
            __syncwarp();
            if (is_last) {
                __threadfence();
                double s = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32) s += ld_cg_f64(P.partials + b);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) {
                    *P.dot_out = s;
                    *P.ticket = 0u;
                }
            }
        }
    }
}

// -----------------------------------------------------------------------------------------------
// variant table
// -----------------------------------------------------------------------------------------------
struct StreamVariant {
    int id, t_nnz, t_rows, stages, ncw, minb;
};
//                 id  T_NNZ T_ROWS STAGES NCW MINB
#define NSK_STREAM_VARIANTS(X) \
    X(0, 4096, 512, 4, 16, 1)  \
    X(1, 2048, 256, 4, 8, 2)   \
    X(2, 4096, 512, 2, 16, 2)  \
    X(3, 2048, 256, 2, 8, 4)   \
    X(4, 1024, 128, 3, 4, 5)   \
    X(5, 4096, 512, 3, 16, 1)  \
    X(6, 3584, 512, 5, 16, 1)  \
    X(7, 4096, 256, 4, 8, 1)

static const StreamVariant g_variants[] = {
#define X(id, t, r, s, w, b) {id, t, r, s, w, b},
    NSK_STREAM_VARIANTS(X)
#undef X
};
static const int g_nvariants = sizeof(g_variants) / sizeof(g_variants[0]);

typedef void (*stream_fn)(const StreamParams);

template <int T, int R, int S, int W, int B>
static stream_fn pick_kernel(int kind, int g, bool muladd, int *smem)
{
    if (kind == 0) {
        *smem = stream_smem_bytes<T, R, S, 0>();
        return muladd ? spmv_stream_kernel<T, R, S, W, B, 0, 1, true> : spmv_stream_kernel<T, R, S, W, B, 0, 1, false>;
    }
    if (kind == 1) {
        *smem = stream_smem_bytes<T, R, S, 1>();
        switch (g) {
            case 4: return spmv_stream_kernel<T, R, S, W, B, 1, 4, false>;
            case 8: return spmv_stream_kernel<T, R, S, W, B, 1, 8, false>;
            case 16: return spmv_stream_kernel<T, R, S, W, B, 1, 16, false>;
            default: return spmv_stream_kernel<T, R, S, W, B, 1, 32, false>;
        }
    }
    *smem = stream_smem_bytes<T, R, S, 2>();
    return muladd ? spmv_stream_kernel<T, R, S, W, B, 2, 1, true> : spmv_stream_kernel<T, R, S, W, B, 2, 1, false>;
}

static stream_fn lookup_kernel(int variant, int kind, int g, bool muladd, int *smem)
{
    switch (variant) {
#define X(id, t, r, s, w, b) \
    case id: return pick_kernel<t, r, s, w, b>(kind, g, muladd, smem);
        NSK_STREAM_VARIANTS(X)
#undef X
    }
    return nullptr;
}

static int default_variant(double mean_row)
{
    // measured on B200, 256^3 7-point (profiles/r01_sweep_spmv_c3.txt): two shallow stages with 4 CTAs
    // per SM (32 consumer warps) beat deep rings with one big CTA: 0.249 ms vs 0.322 ms
    (void)mean_row;
    return 3;
}

void nsk_stream_kernel_config(nsk_ctx_t ctx, double mean_row, int *tile_nnz, int *tile_rows)
{
    int v = ctx->opt.stream_variant > 0 ? (int)ctx->opt.stream_variant - 1 : default_variant(mean_row);
    if (v < 0 || v >= g_nvariants) v = 0;
    *tile_nnz = g_variants[v].t_nnz;
    *tile_rows = g_variants[v].t_rows;
}

// -----------------------------------------------------------------------------------------------
// tiling (host)
// -----------------------------------------------------------------------------------------------
static void make_tiles(const int *ptrow, int n, int t_nnz, int t_rows, const std::vector<int> &breaks,
                       std::vector<nsk_tile> &out, int *nlong)
{
    out.clear();
    *nlong = 0;
    int r = 0;
    size_t bi = 0;
    while (r < n) {
        while (bi < breaks.size() && breaks[bi] <= r) bi++;
        const int seg_end = bi < breaks.size() ? std::min(n, breaks[bi]) : n;  // tiles never straddle a break
        const int nz0 = ptrow[r];
        const int lim = std::min(seg_end, r + t_rows);
        int lo = r, hi = lim;  // last row end with nnz <= t_nnz (binary search, ptrow is monotone)
        while (lo < hi) {
            int mid = lo + (hi - lo + 1) / 2;
            if (ptrow[mid] - nz0 <= t_nnz) lo = mid; else hi = mid - 1;
        }
        int r1 = lo;
        if (r1 == r) {  // a single row longer than a stage
            r1 = r + 1;
            (*nlong)++;
        }
        out.push_back(nsk_tile{r, r1 - r, nz0, ptrow[r1]});
        r = r1;
    }
}

extern std::vector<int> &nsk_csr_host_ptrow(nsk_csr_t A);  // csr.cu

int nsk_get_tiling(nsk_csr_t A, int t_nnz, int t_rows, const nsk_tiling **out)
{
    for (nsk_tiling *T : A->tilings)
        if (T->tile_nnz == t_nnz && T->tile_rows == t_rows) {
            *out = T;
            return NSK_OK;
        }
    nsk_ctx_t ctx = A->ctx;
    nsk_tiling *T = new nsk_tiling();
    make_tiles(nsk_csr_host_ptrow(A).data(), A->n, t_nnz, t_rows, A->breaks, T->h_tiles, &T->nlong);
    T->tile_nnz = t_nnz;
    T->tile_rows = t_rows;
    T->ntiles = (int)T->h_tiles.size();
    size_t bytes = sizeof(nsk_tile) * (size_t)std::max(T->ntiles, 1);
    cudaError_t e = cudaMalloc(&T->d_tiles, bytes + 64);
    if (e == cudaSuccess && T->ntiles)
        e = cudaMemcpy(T->d_tiles, T->h_tiles.data(), sizeof(nsk_tile) * (size_t)T->ntiles, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        nsk_set_error(ctx, "tile table upload failed: %s", cudaGetErrorString(e));
        if (T->d_tiles) cudaFree(T->d_tiles);
        delete T;
        return NSK_ERR_CUDA;
    }
    A->tilings.push_back(T);
    *out = T;
    return NSK_OK;
}

void nsk_free_tilings(nsk_csr_t A)
{
    for (nsk_tiling *T : A->tilings) {
        if (T->d_tiles) cudaFree(T->d_tiles);
        delete T;
    }
    A->tilings.clear();
}

// -----------------------------------------------------------------------------------------------
// launcher
// -----------------------------------------------------------------------------------------------
static int pick_group(double mean_row)
{
    if (mean_row >= 96) return 32;
    if (mean_row >= 40) return 16;
    if (mean_row >= 20) return 8;
    return 4;
}

int nsk_launch_spmv(nsk_csr_t A, const nsk_spmv_args &a)
{
    nsk_ctx_t ctx = A->ctx;
    const int rb = a.row_begin, re = a.row_end;
    if (re <= rb) {
        if (a.dot_w && a.dot_slot >= 0)
            NSK_CUDA(ctx, cudaMemsetAsync(ctx->d_scalars + a.dot_slot, 0, sizeof(double), ctx->stream));
        return NSK_OK;
    }
    const bool muladd = a.mode == NSK_EXACT_MULADD;
    const bool fast = a.mode == NSK_FAST;
    // measured (tools/probe_longrows.py, profiles/r01_probe_longrows.txt): the thread-per-row chain with batched
    // gathers (KIND 0) beats both the lanes-per-row reduction and the gather-to-shared variant at 15 and at 58
    // nonzeros per row (tet mesh: 5.5 vs 4.3 / 2.1 TB/s); the others only pay when a tile holds few rows
    const bool long_rows = A->mean_row > 96.0;
    int sel = (int)ctx->opt.spmv_kernel;
    // default for operators made of pattern tiles (stencils, regular bands): the staged-coefficient sliced-ELL kernel
    if (sel == 0 && rb == 0 && nsk_sell_uniform(A)) sel = 4;
    if (sel == 4 && rb == 0 && nsk_sell_applicable(A)) {  // sliced-ELL tiles (sell.cu)
        double *out = a.y;
        int s = nsk_sell_run(A, 1, a.x, &out, a.mode, &re, a.dot_w, a.dot_slot);
        if (s != NSK_ERR_UNSUPPORTED) { ctx->last_spmv = 4; return s; }
    }
    if (sel == 4) sel = 0;
    if ((sel == 0 || sel == 3) && rb == 0 && nsk_packed_applicable(A)) {
        double *out = a.y;
        int s = nsk_packed_run(A, 1, a.x, &out, a.mode, &re, a.dot_w, a.dot_slot);
        if (s != NSK_ERR_UNSUPPORTED) { ctx->last_spmv = 3; return s; }
    }
    if (sel == 0 || sel == 3) sel = 2;
    ctx->last_spmv = sel;

    if (sel == 1) {
        NSK_REQUIRE(ctx, a.dot_w == nullptr, "fused dot needs the streaming kernel");
        const int rows = re - rb;
        if (fast && long_rows) {
            const int g = pick_group(A->mean_row);
            const long long threads = (long long)rows * g;
            const int blocks = (int)((threads + 255) / 256);
            switch (g) {
                case 4: spmv_vector_kernel<4><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re); break;
                case 8: spmv_vector_kernel<8><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re); break;
                case 16: spmv_vector_kernel<16><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re); break;
                default: spmv_vector_kernel<32><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re); break;
            }
        } else {
            const int blocks = (rows + 255) / 256;
            if (muladd)
                spmv_scalar_kernel<true><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re);
            else
                spmv_scalar_kernel<false><<<blocks, 256, 0, ctx->stream>>>(A->d_ptrow, A->d_indcol, A->d_coef, a.x, a.y, rb, re);
        }
        ctx->launches++;
        NSK_CUDA(ctx, cudaGetLastError());
        return NSK_OK;
    }

    // streaming kernel
    int variant = ctx->opt.stream_variant > 0 ? (int)ctx->opt.stream_variant - 1 : default_variant(A->mean_row);
    if (variant < 0 || variant >= g_nvariants) variant = 0;
    int kind = 0, g = 1;
    if (long_rows) {
        if (fast) { kind = 1; g = pick_group(A->mean_row); }
        else kind = 2;
    }
    if (!fast && ctx->opt.stream_exact_kind == 1) kind = 0;  // experiment switch: thread-per-row chain from global
    if (!fast && ctx->opt.stream_exact_kind == 2) kind = 2;  // ... or gather-to-shared first
    int smem = 0;
    stream_fn fn = lookup_kernel(variant, kind, g, muladd, &smem);
    NSK_REQUIRE(ctx, fn != nullptr, "no such stream kernel variant");
    if (smem > (int)ctx->prop.sharedMemPerBlockOptin) {
        // the gathered-x buffer of KIND 2 does not fit beside 4 big stages: use the 3-stage geometry
        variant = 5;
        fn = lookup_kernel(variant, kind, g, muladd, &smem);
    }
    const StreamVariant &V = g_variants[variant];
    const nsk_tiling *Tp = nullptr;
    NSK_TRY(nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp));
    const nsk_tiling &T = *Tp;

    // tiles covering [rb, re)
    auto cmp = [](const nsk_tile &t, int row) { return t.row0 + t.nrows <= row; };
    int tb = (int)(std::lower_bound(T.h_tiles.begin(), T.h_tiles.end(), rb, cmp) - T.h_tiles.begin());
    int te = (int)(std::lower_bound(T.h_tiles.begin(), T.h_tiles.end(), re, [](const nsk_tile &t, int row) { return t.row0 < row; }) - T.h_tiles.begin());
    if (te <= tb) return NSK_OK;

    static_assert(sizeof(nsk_tile) == 16, "tile descriptor is one 16-byte load");
    NSK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = ctx->opt.spmv_ctas_per_sm > 0 ? (int)ctx->opt.spmv_ctas_per_sm : V.minb;
    int grid = std::min(te - tb, ctx->prop.multiProcessorCount * per_sm);
    if (a.dot_w) NSK_REQUIRE(ctx, grid <= NSK_MAX_PARTIALS, "grid too large for the fused reduction");

    StreamParams P;
    P.tiles = T.d_tiles;
    P.tile_begin = tb;
    P.tile_end = te;
    P.ptrow = A->d_ptrow;
    P.indcol = A->d_indcol;
    P.coef = A->d_coef;
    P.x = a.x;
    P.y = a.y;
    P.row_begin = rb;
    P.row_end = re;
    P.dot_w = a.dot_w;
    P.partials = ctx->d_partials;
    P.ticket = ctx->d_ticket;
    P.dot_out = a.dot_w ? ctx->d_scalars + a.dot_slot : nullptr;
    fn<<<grid, (V.ncw + 1) * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}
