// sell.cu -- sliced-ELL tiles + ONE persistent kernel for y = A x and the fused matrix powers A x .. A^k x on ANY CSR
// operator (stencils and unstructured FEM alike), with the operator's coefficients going HBM/L2 -> registers directly.
//
// Replaces SpMV_CSR / _OPT / _FMA / _AVX2 (reference mpk/SpMV.cpp:6-85) and the fused SpM2V_CSR* / SpM3V / SpM4V
// (mpk/SpM2V.cpp:80-332, mpk/SpMVmulti0.cpp:132-221).  Per-row nonzero order is never changed, so the exact modes are
// bit-identical to the reference's row loop; only the schedule differs.
//
// Why a second fused kernel.  packed.cu stages whole tiles (coefficients, local columns, x runs) in shared memory by
// bulk copies; ncu + stage-cycle timing (profiles/r01_pk_timing.txt) put its floor at the shared-memory
// data path: ~34 B per nonzero written and re-read, 74 % busy at 0.64 ms for k = 4 on 256^3.  The coefficients are
// used exactly once per level, so here they never touch shared memory:
//
//   * a TILE is 256 consecutive rows, one row per consumer thread, stored slot-major per 32-row slice (SELL-32,
//     no row sorting): lane l of warp w reads entry e of its row at val[(slice_off + e) * 32 + l] -- every warp-level
//     load is one 256-byte line pair, no shared memory, no index arithmetic beyond an add;
//   * PATTERN tiles (format P): when every row of a tile is a sub-pattern of one column pattern (stencils, bands: slot e
//     of row r references column r + rel[e]) no per-entry index is stored at all -- 8 bytes per nonzero + 1 mask byte per
//     row (256^3 7-point: 0.98 GB instead of CSR's 1.47 GB); everything else keeps explicit 32-bit columns (format E);
//   * x is gathered with ordinary cached loads: after the dependency warp's acquire (which also invalidates this SM's
//     L1) the level vector below is read through L1, so the +-1 neighbours and the tiles of one chunk share lines;
//   * matrix powers = the level pipeline of packed.cu (teams of CTAs per level, per-group completion counters, window
//     back-pressure sized from the L2 budget) with all device-scope latency moved into two helper warps: a DEPENDENCY
//     warp that polls up to four items ahead, fetches the items' tile descriptors into a shared-memory ring and
//     prefetches level 0's blobs into L2 (cp.async.bulk.prefetch.L2), and a PUBLISHER warp that fences once for
//     everything finished at that moment.  The eight consumer warps are decoupled from each other (a warp only waits
//     for its item's mbarrier) and execute nothing but loads, the fma chain and stores;
//   * completion counters are monotone over launches (an item is complete when its group's counter reaches
//     epoch * group size), so no memset precedes a launch; waits are bounded by the global timer and raise an error
//     flag instead of trapping; k > 1 is launched cooperatively so that co-residency is guaranteed, not assumed.
//
// The host half (tiling, pattern detection, blobs, level schedule, CPU model of the protocol) is plain C++ and is
// tested without a GPU (tests/test_sell_host.py).
#include <algorithm>
#include <array>
#include <atomic>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"
#include "stream_common.cuh"
#include "wave_common.h"

using namespace nskptx;

std::vector<int> &nsk_csr_host_ptrow(nsk_csr_t A);

constexpr int SL_ROWS = 256;      // rows per tile
constexpr int SL_CTHREADS = 128;  // consumer threads per CTA (one warpgroup): thread t owns rows t and t + 128 of a tile
constexpr int SL_NCW = SL_CTHREADS / 32;
constexpr int SL_THREADS = 256;   // + the helper warpgroup (dependency warp, publisher warp, two idle warps)
constexpr int SL_PSLOTS = 8;      // pattern format: at most 8 slots per row
constexpr int SL_MAXCHUNK = 4;    // tiles per item (consecutive tiles taken by one CTA in one go)
constexpr int SL_RING = 4;        // items the dependency warp may run ahead of the slowest consumer warp

enum { SL_FMT_PATTERN = 0, SL_FMT_EXPLICIT = 1 };

struct SlTile {          // 64 bytes = 16 words, fetched by the dependency warp into shared memory
    long long off;       // w0,1  byte offset of the blob (128-byte aligned)
    int row0, nrows;     // w2,3
    int width;           // w4    slots (longest row of the tile)
    int fmt;             // w5
    int rp;              // w6    rows padded to a multiple of 32
    int bytes;           // w7    blob bytes (multiple of 128)
    int rel[8];          // w8..15  format P: column of slot e of local row r = row0 + r + rel[e]
                         //         format E: slots of slice s (its rows' longest), slices stored one after the other
};
static_assert(sizeof(SlTile) == 64, "SlTile is 16 words");

struct SlItem {          // 32 bytes; item i of a level = tiles [i * chunk, (i + 1) * chunk) of the level's tile list
    int group;           // completion group the item reports to
    int ghi;             // forward: groups [0, ghi] of level l-1 complete (-1: none)
    int gback;           // back-pressure: groups [0, gback] of the level that holds this one back (-1: none)
    int pf_bytes;        // blobs of the item's tiles: contiguous bytes starting at pf_off (0: not contiguous)
    long long pf_off;
    int pad[2];
};
static_assert(sizeof(SlItem) == 32, "SlItem is 8 words");

struct SlParams {
    const SlItem *items[NSK_MAX_K];   // per level, in position order
    const SlTile *ltiles[NSK_MAX_K];  // per level: the level's tiles in position order (aliases when all levels agree)
    int count[NSK_MAX_K];             // items per level
    int ntl[NSK_MAX_K];               // tiles per level
    int team[NSK_MAX_K];              // CTAs per level
    int level_rows[NSK_MAX_K];
    double *levels[NSK_MAX_K];
    double *levels2[NSK_MAX_K];
    const double *x, *x2;
    const unsigned char *blobs;
    int *counters;           // [k][ngroups], monotone over launches
    const int *group_size;   // [k][ngroups]
    const int2 *cta_role;    // [grid] {level, index within the level's team}
    int ngroups, n_cols, k, chunk;
    int epoch;               // an item group is complete when counter >= epoch * group size
    int bp_level;            // the level whose progress holds level 0 back (k - 1), -1: no back-pressure
    int muladd;              // 1: NSK_EXACT_MULADD (product and sum rounded separately), 0: fma
    int flags;               // 1: evict-first / streaming hints for data nobody re-reads; 2: L2 prefetch of level 0's blobs;
                             // 4: inner levels load tiles with L1 allocation; 8: ... with an evict-last L2 policy
    int pf_dist;             // items ahead the L2 prefetch runs
    int gpat;                // 1: every tile is stored with the ONE column pattern grel[] (slot e of row r: column r + grel[e])
    int grel[8];
    long long goff8[8];      // 8 * grel[e]: byte offsets, added to a row's x address straight from the constant bank
    int *error;              // host-mapped: set to 1 when a bounded wait expired (protocol bug or a lost CTA)
    long long *timing;       // option pk_timing: per CTA {ns waiting for back-pressure, forward dependencies, ring slot, total, items}
    // fused dot <dot_w, levels[0]> (k = 1 only; CG: p.Ap)
    const double *dot_w;
    double *partials;
    unsigned int *ticket;
    double *dot_out;
};

__host__ __device__ constexpr int sl_round_up(int v, int m) { return (v + m - 1) / m * m; }

// blob layouts (all sections 128-byte aligned)
//   format P:  mask u8[256] | val f64[width][256]      (fixed row stride: slot e of row r at byte 256 + 2048 e + 8 r,
//                                                       so a thread's loads differ by immediates only)
//   format E:  len u16[rp] | col i32[tot][32] | val f64[tot][32]      tot = sum of the slices' widths
static inline int sl_blob_bytes_pattern(int width) { return SL_ROWS + 8 * width * SL_ROWS; }
static inline int sl_blob_bytes_explicit(int rp, int tot) { return sl_round_up(2 * rp, 128) + sl_round_up(4 * 32 * tot, 128) + 8 * 32 * tot; }

__device__ __forceinline__ unsigned long long sl_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// L2 eviction priority of the coefficient / column stream: normal for tiles the next levels re-read, evict-first when
// this level is the last reader of a tile in the launch.  (Measured on 256^3, k = 4: loads marked L1::no_allocate are
// dropped from L2 first as well -- 3.6 GB of HBM reads instead of 1.6 -- so the stream allocates in L1 like any load.)
__device__ __forceinline__ uint64_t sl_policy(bool evict_first)
{
    uint64_t pol;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double sl_ld_coef(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int sl_ld_col(const int *p, uint64_t pol)
{
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// x: ordinary cached load (L1 + L2).  Written as asm so that it can neither become a read-only-path load (the level
// vectors are written by other CTAs of this launch) nor move above the mbarrier wait that orders it.
__device__ __forceinline__ double sl_ld_x(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sl_prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Register re-allocation between the warpgroups of a CTA (sm_90a+): the helper warpgroup hands its registers to the
// consumers, which keep the loads of several rows per thread in flight.
template <int N>
__device__ __forceinline__ void sl_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void sl_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

constexpr unsigned long long SL_TIMEOUT_NS = 4000000000ull;  // bounded waits: 4 s, then the error flag

// Advances the watermark w (all groups < w complete) until it passes `upto`; lanes poll 32 groups at a time.
// Returns the new watermark, or -1 when the wait expired.
__device__ __forceinline__ int sl_wait_groups(const int *cnt, const int *need, int epoch, int ngroups, int w, int upto, int lane)
{
    unsigned long long t0 = 0;
    while (w <= upto) {
        const int g = w + lane;
        bool ok = true;
        if (g < ngroups) ok = ld_acquire_gpu(cnt + g) >= __ldg(need + g) * epoch;
        const unsigned int bad = __ballot_sync(0xffffffffu, !ok);
        const int adv = bad ? __ffs(bad) - 1 : 32;
        w = min(w + adv, ngroups);
        if (w > upto || w >= ngroups) break;
        if (adv == 0) {
            __nanosleep(40);
            const unsigned long long t = sl_now();
            if (t0 == 0) t0 = t;
            else if (t - t0 > SL_TIMEOUT_NS) return -1;
        }
    }
    return w;
}

// What the consumers of a CTA need once per item, kept in shared memory (read back by broadcast LDS).
struct SlCta {
    const double *src[2];
    double *dst[2];
    int row_end, last, muladd, pad;
    int grel[8];  // the operator's global pattern (for code that cannot reach the kernel parameters)
};

template <int NV>
__device__ __forceinline__ void sl_store_row(const SlParams &P, const SlCta &C, int row, double acc0, double acc1, double &dot_acc)
{
    double *dst = C.dst[0];
    if (C.last) __stcs(dst + row, acc0);  // nobody in this launch re-reads the last level
    else dst[row] = acc0;
    if (NV == 2) {
        double *dst2 = C.dst[1];
        if (C.last) __stcs(dst2 + row, acc1);
        else dst2[row] = acc1;
    }
    if (NV == 1 && P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc0, dot_acc);
}

// T pattern tiles of W slots at once: thread t owns rows t and t + 128 of every tile, i.e. 2 T rows, and issues ALL their
// loads (W coefficients + W x entries each) before the first link of any chain -- one memory round trip per item, with
// 2 T W (NV + 1) loads in flight per thread.  d = the item's tile descriptors in shared memory (the tiles share one
// pattern), msk = the rows' slot masks (staged by the dependency warp).  Every load is unconditional so that every
// destination register is defined exactly once: a slot the row lacks reads the row's own x entry and is not used.
template <int NV, int W, int T>
__device__ __forceinline__ void sl_rows_pattern(const int *d, const unsigned char (*msk)[SL_ROWS], const SlParams &P, const SlCta &C,
                                                int t, uint64_t pol, double &dot_acc)
{
    constexpr int WW = W > 0 ? W : 1;
    const int4 r0 = *reinterpret_cast<const int4 *>(d + 8), r1 = *reinterpret_cast<const int4 *>(d + 12);
    const int rel[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const double *src = C.src[0];
    const double *src2 = NV == 2 ? C.src[1] : nullptr;
    const int cmax = P.n_cols - 1;
    double a[T][2][WW], xv[NV][T][2][WW];
    unsigned int m[T][2];
    int row[T][2];
    bool live[T][2];
#pragma unroll
    for (int j = 0; j < T; j++) {
        const int4 q = *reinterpret_cast<const int4 *>(d + 16 * j);  // off lo, off hi, row0, nrows
        const long long off = ((long long)(unsigned int)q.x) | ((long long)q.y << 32);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = t + SL_CTHREADS * h;
            m[j][h] = msk[j][r];
            row[j][h] = q.z + r;
            live[j][h] = r < q.w && row[j][h] < C.row_end;
            const double *val = reinterpret_cast<const double *>(P.blobs + off + SL_ROWS) + r;
            const int rowc = min(row[j][h], cmax);  // padding rows of a short tile stay inside x
#pragma unroll
            for (int e = 0; e < W; e++) {
                a[j][h][e] = sl_ld_coef(val + e * SL_ROWS, pol);
                const int idx = rowc + (rel[e] & -(int)((m[j][h] >> e) & 1u));
                xv[0][j][h][e] = sl_ld_x(src + idx);
                if (NV == 2) xv[NV - 1][j][h][e] = sl_ld_x(src2 + idx);
            }
        }
    }
    double acc0[T][2], acc1[T][2];
    if (C.muladd) {
#pragma unroll
        for (int j = 0; j < T; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                acc0[j][h] = 0.0;
                acc1[j][h] = 0.0;
#pragma unroll
                for (int e = 0; e < W; e++)
                    if (m[j][h] & (1u << e)) {
                        acc0[j][h] = row_op<true>(a[j][h][e], xv[0][j][h][e], acc0[j][h]);
                        if (NV == 2) acc1[j][h] = row_op<true>(a[j][h][e], xv[NV - 1][j][h][e], acc1[j][h]);
                    }
            }
    } else {
#pragma unroll
        for (int j = 0; j < T; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                acc0[j][h] = 0.0;
                acc1[j][h] = 0.0;
#pragma unroll
                for (int e = 0; e < W; e++)
                    if (m[j][h] & (1u << e)) {
                        acc0[j][h] = row_op<false>(a[j][h][e], xv[0][j][h][e], acc0[j][h]);
                        if (NV == 2) acc1[j][h] = row_op<false>(a[j][h][e], xv[NV - 1][j][h][e], acc1[j][h]);
                    }
            }
    }
#pragma unroll
    for (int j = 0; j < T; j++)
#pragma unroll
        for (int h = 0; h < 2; h++)
            if (live[j][h]) sl_store_row<NV>(P, C, row[j][h], acc0[j][h], acc1[j][h], dot_acc);
}

template <int NV, int T>
__device__ __forceinline__ void sl_rows_pattern_any(int width, const int *d, const unsigned char (*msk)[SL_ROWS], const SlParams &P,
                                                    const SlCta &C, int t, uint64_t pol, double &dot_acc)
{
    switch (width) {  // uniform: one straight-line instance per width
    case 0: sl_rows_pattern<NV, 0, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 1: sl_rows_pattern<NV, 1, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 2: sl_rows_pattern<NV, 2, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 3: sl_rows_pattern<NV, 3, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 4: sl_rows_pattern<NV, 4, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 5: sl_rows_pattern<NV, 5, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 6: sl_rows_pattern<NV, 6, T>(d, msk, P, C, t, pol, dot_acc); break;
    case 7: sl_rows_pattern<NV, 7, T>(d, msk, P, C, t, pol, dot_acc); break;
    default:  // 8 slots: the widest pattern keeps one tile less in flight (registers)
        if (T >= 2 && (NV == 2 || T >= 3)) {
            sl_rows_pattern<NV, 8, (T > 1 ? T - 1 : 1)>(d, msk, P, C, t, pol, dot_acc);
            sl_rows_pattern<NV, 8, 1>(d + 16 * (T - 1), msk + (T - 1), P, C, t, pol, dot_acc);
        } else {
            sl_rows_pattern<NV, 8, T>(d, msk, P, C, t, pol, dot_acc);
        }
        break;
    }
}

// One tile, format E (explicit 32-bit columns, per-slice widths): thread t owns rows t and t + 128; EB entries of both
// rows per round trip (columns and coefficients, then the gathers).
template <int NV, int EB>
__device__ __forceinline__ void sl_tile_explicit(const int *d, const SlParams &P, const SlCta &C, int t, uint64_t pol, double &dot_acc)
{
    const double *src = C.src[0];
    const double *src2 = NV == 2 ? C.src[1] : nullptr;
    const long long off = ((long long)(unsigned int)d[0]) | ((long long)d[1] << 32);
    const int row0 = d[2], nrows = d[3], rp = d[6];
    const unsigned char *blob = P.blobs + off;
    const int lane = t & 31, w = t >> 5;  // rows t and t + 128 lie in slices w and w + 4
    int soff[2] = {0, 0}, tot = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int ws = d[8 + s];
        if (s < w) soff[0] += ws;
        if (s < w + 4) soff[1] += ws;
        tot += ws;
    }
    const unsigned char *colbase = blob + sl_round_up(2 * rp, 128);
    const unsigned char *valbase = colbase + sl_round_up(128 * tot, 128);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = t + SL_CTHREADS * h;
        if ((r & ~31) >= rp) continue;  // the tile has no slice for this warp's second row
        const int mine = d[8 + w + 4 * h];
        const int len = (int)__ldg(reinterpret_cast<const unsigned short *>(blob) + r);
        const int *col = reinterpret_cast<const int *>(colbase) + (size_t)soff[h] * 32 + lane;
        const double *val = reinterpret_cast<const double *>(valbase) + (size_t)soff[h] * 32 + lane;
        double acc0 = 0.0, acc1 = 0.0;
        for (int e0 = 0; e0 < mine; e0 += EB) {
            int cc[EB];
            double a[EB], xv[NV][EB];
            // every load unconditional (each destination register defined once): a batch that runs past the slice's
            // last slot re-reads that slot; padding entries carry a valid column (0); neither is used
#pragma unroll
            for (int u = 0; u < EB; u++) {
                const int e = min(e0 + u, mine - 1);
                cc[u] = sl_ld_col(col + (size_t)e * 32, pol);
                a[u] = sl_ld_coef(val + (size_t)e * 32, pol);
            }
#pragma unroll
            for (int u = 0; u < EB; u++) {
                xv[0][u] = sl_ld_x(src + cc[u]);
                if (NV == 2) xv[NV - 1][u] = sl_ld_x(src2 + cc[u]);
            }
#pragma unroll
            for (int u = 0; u < EB; u++)
                if (e0 + u < len) {
                    if (C.muladd) {
                        acc0 = row_op<true>(a[u], xv[0][u], acc0);
                        if (NV == 2) acc1 = row_op<true>(a[u], xv[NV - 1][u], acc1);
                    } else {
                        acc0 = row_op<false>(a[u], xv[0][u], acc0);
                        if (NV == 2) acc1 = row_op<false>(a[u], xv[NV - 1][u], acc1);
                    }
                }
        }
        const int row = row0 + r;
        if (r < nrows && row < C.row_end) sl_store_row<NV>(P, C, row, acc0, acc1, dot_acc);
    }
}

// ---- streaming consumers (operators made of pattern tiles of one width W) ---------------------------------------
// A consumer thread keeps D tiles in flight (rows t and t + 128 of each) in D statically named register sets and
// software-pipelines the tile stream of its CTA: [issue the loads of tile n + D - 1] [chain + store tile n] ... so the
// load pipe is fed continuously instead of in bursts, and a load has D - 1 tiles' worth of work to complete.  The sets
// rotate in a fixed order (fill order = compute order).  A tile whose item is not ready yet is NOT waited for while
// other sets hold tiles (a bubble passes through instead): finishing an item never depends on a later item's inputs,
// which is what the level schedule's deadlock-freedom argument (and its CPU model) assumes.
template <int NV, int W, int R>
struct SlSet {
    double a[R][W], xv[NV][R][W];
    unsigned int m[R];
    int row[R];
    int state;  // bit 0: holds a tile; bit 1: last tile of its item; bits 2, 3: row h is stored; bits 8..: ring slot
};

struct SlCursor {
    int it, j, ntile;  // next tile: j-th of item it; ntile < 0: the item has not been opened yet
};

// Fills `S` with the next tile of the stream; returns false when the stream has ended or (may_block == false) its item
// is not ready.
template <int NV, int W, int R>
__device__ __forceinline__ bool sl_stream_issue(SlSet<NV, W, R> &S, SlCursor &cur, int n_my, bool may_block, const SlParams &P,
                                                const SlCta &C, uint64_t *s_ready, const int (*s_hdr)[2],
                                                const int (*s_desc)[SL_MAXCHUNK * 16],
                                                const unsigned char (*s_mask)[SL_MAXCHUNK][SL_ROWS], int t, uint64_t pol)
{
    S.state = 0;
    if (cur.it >= n_my) return false;
    const int slot = cur.it % SL_RING;
    if (cur.ntile < 0) {
        const uint32_t parity = (uint32_t)(cur.it / SL_RING) & 1u;
        if (!mbar_try_wait(&s_ready[slot], parity)) {
            if (!may_block) return false;
            mbar_wait(&s_ready[slot], parity);
        }
        cur.ntile = s_hdr[slot][0];
    }
    const int *d = s_desc[slot] + 16 * cur.j;
    const int4 q = *reinterpret_cast<const int4 *>(d);  // off lo, off hi, row0, nrows
    const int4 r0 = *reinterpret_cast<const int4 *>(d + 8), r1 = *reinterpret_cast<const int4 *>(d + 12);
    const int rel[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const long long off = ((long long)(unsigned int)q.x) | ((long long)q.y << 32);
    const double *src = C.src[0];
    const double *src2 = NV == 2 ? C.src[1] : nullptr;
    const int cmax = P.n_cols - 1;
    int state = 1 | (slot << 8);
#pragma unroll
    for (int h = 0; h < R; h++) {
        const int r = t + (SL_ROWS / R) * h;
        S.m[h] = s_mask[slot][cur.j][r];
        S.row[h] = q.z + r;
        if (r < q.w && S.row[h] < C.row_end) state |= 4 << h;
        const double *val = reinterpret_cast<const double *>(P.blobs + off + SL_ROWS) + r;
        const int rowc = min(S.row[h], cmax);  // padding rows of a short tile stay inside x
#pragma unroll
        for (int e = 0; e < W; e++) {
            S.a[h][e] = sl_ld_coef(val + e * SL_ROWS, pol);
            const int idx = rowc + (rel[e] & -(int)((S.m[h] >> e) & 1u));  // a slot the row lacks reads its own x entry
            S.xv[0][h][e] = sl_ld_x(src + idx);
            if (NV == 2) S.xv[NV - 1][h][e] = sl_ld_x(src2 + idx);
        }
    }
    if (++cur.j == cur.ntile) {
        state |= 2;
        cur.j = 0;
        cur.ntile = -1;
        cur.it++;
    }
    S.state = state;
    return true;
}

template <int NV, int W, int R>
__device__ __forceinline__ void sl_stream_compute(const SlSet<NV, W, R> &S, const SlParams &P, const SlCta &C, unsigned int *s_fin,
                                                  double &dot_acc)
{
    double acc0[R], acc1[R];
#pragma unroll
    for (int h = 0; h < R; h++) {
        acc0[h] = 0.0;
        acc1[h] = 0.0;
        if (C.muladd) {
#pragma unroll
            for (int e = 0; e < W; e++)
                if (S.m[h] & (1u << e)) {
                    acc0[h] = row_op<true>(S.a[h][e], S.xv[0][h][e], acc0[h]);
                    if (NV == 2) acc1[h] = row_op<true>(S.a[h][e], S.xv[NV - 1][h][e], acc1[h]);
                }
        } else {
#pragma unroll
            for (int e = 0; e < W; e++)
                if (S.m[h] & (1u << e)) {
                    acc0[h] = row_op<false>(S.a[h][e], S.xv[0][h][e], acc0[h]);
                    if (NV == 2) acc1[h] = row_op<false>(S.a[h][e], S.xv[NV - 1][h][e], acc1[h]);
                }
        }
    }
#pragma unroll
    for (int h = 0; h < R; h++)
        if (S.state & (4 << h)) sl_store_row<NV>(P, C, S.row[h], acc0[h], acc1[h], dot_acc);
    if (S.state & 2) {  // last tile of its item: the slot may be reused and the item published
        __syncwarp();
        if ((threadIdx.x & 31) == 0) red_release_cta_shared_add(&s_fin[S.state >> 8], 1u);
    }
}

template <int NV, int W, int D, int R>
__device__ __forceinline__ double sl_consume_stream(const SlParams &P, const SlCta &C, int n_my, uint64_t *s_ready, unsigned int *s_fin,
                                                    const int (*s_hdr)[2], const int (*s_desc)[SL_MAXCHUNK * 16],
                                                    const unsigned char (*s_mask)[SL_MAXCHUNK][SL_ROWS], uint64_t pol)
{
    const int t = threadIdx.x;
    SlSet<NV, W, R> S[D];
    SlCursor cur{0, 0, -1};
    double dot_acc = 0.0;
    int pending = 0;
#pragma unroll
    for (int d = 0; d < D - 1; d++)
        pending += sl_stream_issue<NV, W, R>(S[d], cur, n_my, pending == 0, P, C, s_ready, s_hdr, s_desc, s_mask, t, pol) ? 1 : 0;
    while (cur.it < n_my || pending > 0) {
#pragma unroll
        for (int d = 0; d < D; d++) {
            // the set computed in the previous step is free: fill it (or let a bubble through), then compute the oldest
            pending += sl_stream_issue<NV, W, R>(S[(d + D - 1) % D], cur, n_my, pending == 0, P, C, s_ready, s_hdr, s_desc, s_mask, t,
                                              pol) ? 1 : 0;
            if (S[d].state & 1) {
                sl_stream_compute<NV, W, R>(S[d], P, C, s_fin, dot_acc);
                S[d].state = 0;
                pending--;
            }
        }
    }
    return dot_acc;
}

// CTA = two warpgroups.  Warps 0..3: consumers (rows t and t + 128 of every tile; most of the CTA's registers).
// Warp 4: dependency warp.  Warp 5: publisher.  Warps 6, 7 only give their registers away.
// PW = 0: any operator, an item at a time (T tiles of one pattern at once, everything else tile by tile).
// PW > 0: operators made of pattern tiles of width PW only -- streaming consumers with T tiles in flight per thread.
// R = rows of a tile per consumer thread (streaming kernel): 2 -> one consumer warpgroup with 232 registers per thread,
// 1 -> two consumer warpgroups (256 threads, 104 registers): twice the warps, half the loads in flight per warp.
template <int R>
struct SlShape {
    static constexpr int CT = SL_ROWS / R;       // consumer threads
    static constexpr int NCW = CT / 32;          // consumer warps
    static constexpr int NT = CT + 128;          // + the helper warpgroup
    static constexpr int REGS = R == 2 ? 232 : 104;  // consumers after the helper warpgroup has dropped to 24
    static constexpr int LAUNCH_REGS = R == 2 ? 128 : 80;  // what the launch bounds must give for the pool to add up
};

template <int NV, int T, int PW, int R>
__global__ void __launch_bounds__(SlShape<R>::NT, 2) sell_kernel(const SlParams P)
{
    constexpr int NCW = SlShape<R>::NCW;
    static_assert(NV == 1 || NV == 2, "one or two right-hand sides");
    static_assert(T >= 1 && T <= SL_MAXCHUNK, "tiles per item");
    __shared__ __align__(16) int s_desc[SL_RING][SL_MAXCHUNK * 16];               // tile descriptors of the items in flight
    __shared__ __align__(8) unsigned char s_mask[SL_RING][SL_MAXCHUNK][SL_ROWS];  // format P: the rows' slot masks
    __shared__ int s_hdr[SL_RING][2];         // tiles in the item; their common width when they share one pattern, else -1
    __shared__ uint64_t s_ready[SL_RING];     // item's inputs complete + descriptors in place (dependency warp arrives)
    __shared__ unsigned int s_fin[SL_RING];   // consumer warps that finished the slot's item, counted over the whole launch
    __shared__ double s_red[NCW];
    __shared__ SlCta s_cta;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int2 role = __ldg(P.cta_role + blockIdx.x);
    const int level = role.x;
    const int c = role.y;
    const int G = P.team[level];
    const int count = P.count[level];
    const int n_my = c < count ? (count - c + G - 1) / G : 0;
    if (tid == 0) {
        for (int s = 0; s < SL_RING; s++) {
            mbar_init(&s_ready[s], 1);
            s_fin[s] = 0u;
        }
        fence_mbar_init();
        s_cta.src[0] = level == 0 ? P.x : P.levels[level - 1];
        s_cta.src[1] = NV == 2 ? (level == 0 ? P.x2 : P.levels2[level - 1]) : nullptr;
        s_cta.dst[0] = P.levels[level];
        s_cta.dst[1] = NV == 2 ? P.levels2[level] : nullptr;
        s_cta.row_end = P.level_rows[level];
        s_cta.last = ((P.flags & 1) && level == P.k - 1) ? 1 : 0;
        s_cta.muladd = P.muladd;
    }
    __syncthreads();

    if (warp >= NCW) {
        sl_reg_dec<24>();
        if (warp == NCW) {
            // ===== dependency warp: per item, in order -- (1) fetch the tile descriptors into registers, (2) wait until
            // the level below has completed the groups the item reads and the level that holds this one back has
            // advanced, (3) wait for the ring slot (the item four back is finished by every consumer warp), (4) stage
            // descriptors + slot masks and hand over. =====
            const int M = P.chunk;
            const int ntl = P.ntl[level];
            const int *itw = reinterpret_cast<const int *>(P.items[level]);
            const int4 *tl4 = reinterpret_cast<const int4 *>(P.ltiles[level]);
            const bool fwd = level > 0;
            const bool back = level == 0 && P.bp_level > 0;
            const int *cnt_f = P.counters + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
            const int *need_f = P.group_size + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
            const int *cnt_b = P.counters + (size_t)(back ? P.bp_level : 0) * P.ngroups;
            const int *need_b = P.group_size + (size_t)(back ? P.bp_level : 0) * P.ngroups;
            const bool prefetch = (P.flags & 2) && level == 0 && P.pf_dist > 0;
            int wf = 0, wb = 0;
            bool broken = false;
            for (int it = 0; it < n_my; ++it) {
                const long long i = (long long)c + (long long)it * G;
                const int s = it % SL_RING;
                // lanes 0..7: the item's words; lane l: descriptor quarter l % 4 of tile l / 4 (4 int4 per tile)
                int iw = 0;
                if (lane < 8) iw = __ldg(itw + i * 8 + lane);
                const long long t0 = i * M;
                const int ntile = (int)min((long long)M, (long long)ntl - t0);
                int4 dq = make_int4(0, 0, 0, 0);
                if (lane < 4 * ntile) dq = __ldg(tl4 + t0 * 4 + lane);
                if (prefetch && it + P.pf_dist < n_my) {
                    // level 0 streams its tiles from HBM: pull a later item's bytes into L2 now
                    const long long ip = (long long)c + (long long)(it + P.pf_dist) * G;
                    int pw = 0;
                    if (lane >= 3 && lane < 6) pw = __ldg(itw + ip * 8 + lane);  // pf_bytes, pf_off lo / hi
                    const int pb = __shfl_sync(0xffffffffu, pw, 3);
                    const unsigned int lo = (unsigned int)__shfl_sync(0xffffffffu, pw, 4);
                    const int hi = __shfl_sync(0xffffffffu, pw, 5);
                    if (lane == 0 && pb > 0) sl_prefetch_l2(P.blobs + (((long long)hi << 32) | lo), (uint32_t)pb);
                }
                // do the item's tiles share one pattern (format P, same width, same offsets)?
                bool same = true;
                {
                    const int q = lane & 3;
                    const int rx = __shfl_sync(0xffffffffu, dq.x, q), ry = __shfl_sync(0xffffffffu, dq.y, q);
                    const int rz = __shfl_sync(0xffffffffu, dq.z, q), rw = __shfl_sync(0xffffffffu, dq.w, q);
                    if (lane < 4 * ntile) {
                        if (q == 1) same = dq.x == rx && dq.y == ry && dq.y == SL_FMT_PATTERN;
                        if (q >= 2) same = dq.x == rx && dq.y == ry && dq.z == rz && dq.w == rw;
                    }
                    same = __all_sync(0xffffffffu, same);
                }
                const int width0 = __shfl_sync(0xffffffffu, dq.x, 1);
                const int ghi = __shfl_sync(0xffffffffu, iw, 1);
                const int gback = __shfl_sync(0xffffffffu, iw, 2);
                if (!broken) {
                    if (back && gback >= wb) wb = sl_wait_groups(cnt_b, need_b, P.epoch, P.ngroups, wb, gback, lane);
                    if (fwd && ghi >= wf && wb >= 0) wf = sl_wait_groups(cnt_f, need_f, P.epoch, P.ngroups, wf, ghi, lane);
                    if (wb < 0 || wf < 0) broken = true;
                }
                if (it >= SL_RING && !broken) {
                    const unsigned int need = (unsigned int)NCW * (unsigned int)(it / SL_RING);
                    unsigned long long t1 = 0;
                    while (ld_acquire_cta_shared_u32(&s_fin[s]) < need) {
                        __nanosleep(20);
                        const unsigned long long t = sl_now();
                        if (t1 == 0) t1 = t;
                        else if (t - t1 > SL_TIMEOUT_NS) { broken = true; break; }
                    }
                }
                if (broken && lane == 0) *P.error = 1;  // keep going without waiting: the launch ends, the host reports it
                // slot masks of the item's pattern tiles (256 bytes each: 8 per lane); their addresses come from the descriptors
                for (int j = 0; j < ntile; j++) {
                    const unsigned int lo = (unsigned int)__shfl_sync(0xffffffffu, dq.x, 4 * j);
                    const int hi = __shfl_sync(0xffffffffu, dq.y, 4 * j);
                    const int fmt = __shfl_sync(0xffffffffu, dq.y, 4 * j + 1);
                    uint2 mk = make_uint2(0u, 0u);
                    if (fmt == SL_FMT_PATTERN) mk = __ldg(reinterpret_cast<const uint2 *>(P.blobs + (((long long)hi << 32) | lo)) + lane);
                    reinterpret_cast<uint2 *>(s_mask[s][j])[lane] = mk;
                }
                if (lane < 4 * SL_MAXCHUNK) reinterpret_cast<int4 *>(s_desc[s])[lane] = dq;
                if (lane == 0) {
                    s_hdr[s][0] = ntile;
                    s_hdr[s][1] = same ? width0 : -1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_ready[s]);  // release.cta: descriptors + everything acquired above
            }
        } else if (warp == NCW + 1 && P.k > 1) {
            // ===== publisher: one gpu-scope fence for everything found finished at that moment, then one RED per item.
            // Consumers bump s_fin[slot] with release.cta after their stores; the acquire here + fence + RED is
            // cumulative over those stores. =====
            const int *itw = reinterpret_cast<const int *>(P.items[level]);
            int *cnt = P.counters + (size_t)level * P.ngroups;
            int grp = 0;  // lane u: group of item it0 + u
            for (int it = 0; it < n_my;) {
                const int j = it & 31;
                if (j == 0) {
                    grp = 0;
                    if (it + lane < n_my) grp = __ldg(itw + ((long long)c + (long long)(it + lane) * G) * 8);
                }
                int n = 0;
                unsigned long long t0 = 0;
                bool broken = false;
                for (;;) {  // every consecutive finished item (at most to the end of this batch of 32)
                    const int i2 = it + n;
                    bool ok = false;
                    if (i2 < n_my && (n == 0 || (i2 & 31) != 0) && n < SL_RING) {
                        const unsigned int need = (unsigned int)NCW * (unsigned int)(i2 / SL_RING + 1);
                        ok = ld_acquire_cta_shared_u32(&s_fin[i2 % SL_RING]) >= need;
                    }
                    ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
                    if (ok) { ++n; continue; }
                    if (n > 0) break;
                    __nanosleep(20);
                    const unsigned long long t = sl_now();
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > SL_TIMEOUT_NS) { broken = true; break; }
                }
                if (broken) {
                    if (lane == 0) *P.error = 1;
                    break;
                }
                if (lane == 0) __threadfence();
                __syncwarp();
                for (int u = 0; u < n; u++) {
                    const int g = __shfl_sync(0xffffffffu, grp, (it + u) & 31);
                    if (lane == 0) red_relaxed_gpu_add(cnt + g, 1);
                }
                it += n;
            }
        }
        return;
    }

    // ===== consumer warpgroup =====
    sl_reg_inc<SlShape<R>::REGS>();
    const uint64_t pol = sl_policy(s_cta.last != 0);
    double dot_acc = 0.0;
    if (PW > 0) dot_acc = sl_consume_stream<NV, PW, T, R>(P, s_cta, n_my, s_ready, s_fin, s_hdr, s_desc, s_mask, pol);
    for (int it = 0; PW == 0 && it < n_my; ++it) {
        const int s = it % SL_RING;
        mbar_wait(&s_ready[s], (it / SL_RING) & 1);
        const int ntile = s_hdr[s][0], width = s_hdr[s][1];
        if (ntile == T && width >= 0) {
            sl_rows_pattern_any<NV, T>(width, s_desc[s], s_mask[s], P, s_cta, tid, pol, dot_acc);
        } else {
            for (int j = 0; j < ntile; j++) {  // a short last item, mixed formats or explicit tiles: tile by tile
                const int *d = s_desc[s] + 16 * j;
                if (d[5] == SL_FMT_PATTERN) sl_rows_pattern_any<NV, 1>(d[4], d, s_mask[s] + j, P, s_cta, tid, pol, dot_acc);
                else sl_tile_explicit<NV, NV == 2 ? 8 : 16>(d, P, s_cta, tid, pol, dot_acc);
            }
        }
        __syncwarp();
        if (lane == 0) red_release_cta_shared_add(&s_fin[s], 1u);  // the slot may be reused and the item published
    }

    if (P.dot_w) {
        // deterministic: lanes -> warp (xor tree), warps in order, CTAs in order (last CTA finishes)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_xor_sync(0xffffffffu, dot_acc, o);
        if (lane == 0) s_red[warp] = dot_acc;
        named_bar_sync(2, SlShape<R>::CT);
        if (warp == 0) {
            __shared__ bool is_last;
            if (lane == 0) {
                double sum = 0.0;
                for (int w = 0; w < NCW; w++) sum += s_red[w];
                P.partials[blockIdx.x] = sum;
                __threadfence();
                const unsigned int ticket = atomicAdd(P.ticket, 1u);
                is_last = (ticket == gridDim.x - 1);
            }
            __syncwarp();
            if (is_last) {
                __threadfence();
                double sum = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32) sum += ld_cg_f64(P.partials + b);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) {
                    *P.dot_out = sum;
                    *P.ticket = 0u;
                }
            }
        }
    }
}

// -----------------------------------------------------------------------------------------------
// kernel instances: NV right-hand sides x T tiles per item (2 T rows per consumer thread in flight)
// -----------------------------------------------------------------------------------------------
typedef void (*sl_fn)(const SlParams);
struct SlLaunch {
    sl_fn fn = nullptr;
    int threads = 0, launch_regs = 0;
};
// generic kernel: chunk = tiles per item = tiles of one pattern a consumer thread takes at once
static SlLaunch sl_lookup(int nv, int chunk)
{
    SlLaunch L;
    L.threads = SlShape<2>::NT;
    L.launch_regs = SlShape<2>::LAUNCH_REGS;
    if (nv == 2) L.fn = chunk >= 2 ? sell_kernel<2, 2, 0, 2> : sell_kernel<2, 1, 0, 2>;
    else L.fn = chunk >= 3 ? sell_kernel<1, 3, 0, 2> : chunk == 2 ? sell_kernel<1, 2, 0, 2> : sell_kernel<1, 1, 0, 2>;
    return L;
}
static int sl_max_chunk(int nv) { return nv == 2 ? 2 : 3; }
// streaming kernel (all tiles pattern tiles of width w): depth tiles in flight per thread, rows of a tile per thread
template <int NV, int D, int R>
static sl_fn sl_lookup_stream_w(int w)
{
    switch (w) {
    case 1: return sell_kernel<NV, D, 1, R>;
    case 2: return sell_kernel<NV, D, 2, R>;
    case 3: return sell_kernel<NV, D, 3, R>;
    case 4: return sell_kernel<NV, D, 4, R>;
    case 5: return sell_kernel<NV, D, 5, R>;
    case 6: return sell_kernel<NV, D, 6, R>;
    case 7: return sell_kernel<NV, D, 7, R>;
    case 8: return sell_kernel<NV, D, 8, R>;
    }
    return nullptr;
}
static SlLaunch sl_lookup_stream(int nv, int w, int depth, int rows)
{
    SlLaunch L;
    if (rows == 1) {  // 104 registers: two tiles in flight (one for two right-hand sides of a wide pattern)
        L.threads = SlShape<1>::NT;
        L.launch_regs = SlShape<1>::LAUNCH_REGS;
        if (nv == 2) L.fn = w <= 5 ? sl_lookup_stream_w<2, 2, 1>(w) : sl_lookup_stream_w<2, 1, 1>(w);
        else L.fn = depth >= 2 ? sl_lookup_stream_w<1, 2, 1>(w) : sl_lookup_stream_w<1, 1, 1>(w);
        return L;
    }
    L.threads = SlShape<2>::NT;
    L.launch_regs = SlShape<2>::LAUNCH_REGS;
    if (nv == 2) L.fn = sl_lookup_stream_w<2, 2, 2>(w);
    else {
        if (w == 8) depth = std::min(depth, 2);  // 8 slots x 2 rows x 3 tiles would not fit the consumers' registers
        L.fn = depth >= 3 ? sl_lookup_stream_w<1, 3, 2>(w) : sl_lookup_stream_w<1, 2, 2>(w);
    }
    return L;
}

// ---- staged-coefficient kernel (operators made of pattern tiles of one width W) ------------------------------------
// The coefficients of a tile are ONE contiguous block (mask + W x 256 doubles): a producer lane streams them into a
// shared-memory ring with 1-D bulk copies (TMA, SASS UBLKCP) as soon as a stage is free -- they depend on no other CTA,
// so their HBM / L2 latency is hidden by the ring, not by registers.  Consumer threads (one row each, many warps) only
// hold the row's x entries in flight: W ordinary cached loads issued the moment the item's inputs are complete, then the
// chain with coefficients read from the stage.  Per row and level the shared-memory / L1 data path carries the
// coefficients twice (bulk write + read) and x once -- against coefficients, local columns AND x twice in packed.cu.
constexpr int SL_EXPLICIT_STAGE = 24576;         // stage of the explicit-column instances: lengths + columns of a tile
constexpr int SLT_NCW = SL_ROWS / 32;            // consumer warps: thread t owns row t of every tile
constexpr int SLT_THREADS = SL_ROWS + 64;        // + dependency warp + service warp (producer and publisher)


// ---- consumers of an operator stored with ONE global pattern (staged kernel) ----------------------------------------
// Everything per-thread that does not change over the launch is folded into a few registers (shared-memory addresses as
// 32-bit words, x / y pointers already offset by the thread's row), the pattern's byte offsets come from the constant
// bank, and a warp whose 32 rows have every slot runs straight-line code: W gathers, W shared-memory loads, W links of
// the chain, one store.  Warps with a row that lacks a slot (domain faces) take the masked variant of the same code.
__device__ __forceinline__ bool sl_mbar_try_wait_a(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sl_mbar_wait_a(uint32_t bar, uint32_t parity)
{
    if (sl_mbar_try_wait_a(bar, parity)) return;
    uint32_t spins = 0;
    while (!sl_mbar_try_wait_a(bar, parity))
        if (++spins > (1u << 26)) __trap();  // bounded: a protocol bug must not hang the GPU
}
__device__ __forceinline__ void sl_mbar_arrive_a(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int OFF>
__device__ __forceinline__ double sl_lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
    return v;
}
__device__ __forceinline__ unsigned int sl_lds_u8(uint32_t a)
{
    unsigned int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double sl_ld_x_b(const char *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int NV, int W, int E>
struct SlChain {
    template <bool MULADD>
    static __device__ __forceinline__ void run(uint32_t a_val, const double (&xv)[NV][W], unsigned int m, bool masked, double &acc0, double &acc1)
    {
        if constexpr (E < W) {
            if (!masked || (m & (1u << E))) {
                const double a = sl_lds_f64<E * SL_ROWS * 8>(a_val);
                acc0 = row_op<MULADD>(a, xv[0][E], acc0);
                if (NV == 2) acc1 = row_op<MULADD>(a, xv[NV - 1][E], acc1);
            }
            SlChain<NV, W, E + 1>::template run<MULADD>(a_val, xv, m, masked, acc0, acc1);
        }
    }
};

// x entry at byte address base + off (one 64-bit add the compiler cannot fold into anything else, then the load)
__device__ __forceinline__ double sl_ld_x_off(const char *base, long long off)
{
    double v;
    asm volatile("{\n\t.reg .b64 a;\n\tadd.s64 a, %1, %2;\n\tld.global.f64 %0, [a];\n\t}" : "=d"(v) : "l"(base), "l"(off) : "memory");
    return v;
}
__device__ __forceinline__ const char *sl_add_b(const char *base, long long off)
{
    const char *r;
    asm volatile("add.s64 %0, %1, %2;" : "=l"(r) : "l"(base), "l"(off));
    return r;
}

template <int NV, int W, int NS>
__device__ __forceinline__ void sl_consume_gpat(const SlParams &P, const SlCta &C, int n_my, int tid, const unsigned char *stages,
                                                const uint64_t *s_full, const uint64_t *s_empty, const uint64_t *s_ready,
                                                unsigned int *s_fin, const int (*s_hdr)[2], const int (*s_desc)[SL_MAXCHUNK * 16],
                                                int &it_io, int &st_io, uint32_t &ph_io)
{
    constexpr int STAGE = SL_ROWS + 8 * W * SL_ROWS;
    constexpr unsigned int FULL = (1u << W) - 1u;
    const uint32_t a_full = smem_u32(s_full), a_empty = smem_u32(s_empty), a_ready = smem_u32(s_ready);
    const uint32_t a_mask = smem_u32(stages) + (uint32_t)tid;
    const uint32_t a_val = smem_u32(stages) + SL_ROWS + 8u * (uint32_t)tid;
    const char *const xsrc = reinterpret_cast<const char *>(C.src[0] + tid);
    const char *const xsrc2 = NV == 2 ? reinterpret_cast<const char *>(C.src[1] + tid) : nullptr;
    char *const ydst = reinterpret_cast<char *>(C.dst[0] + tid);
    char *const ydst2 = NV == 2 ? reinterpret_cast<char *>(C.dst[1] + tid) : nullptr;
    const bool last = C.last != 0;
    const bool muladd = C.muladd != 0;
    const bool lane0 = (tid & 31) == 0;
    // byte offset of the next tile's stage, of its barriers; parity of its completion
    uint32_t so = (uint32_t)st_io * STAGE, sb = 8u * (uint32_t)st_io, ph = ph_io;
    int it = it_io;
    for (; it < n_my; ++it) {
        const int slot = it % SL_RING;
        sl_mbar_wait_a(a_ready + 8u * slot, (uint32_t)(it / SL_RING) & 1u);
        const int ntile = s_hdr[slot][0];
        if (s_hdr[slot][1] == 0) break;  // an item with an exception tile (a pattern of its own): the caller's general loop takes it
        const int *d = s_desc[slot] + 2;
        for (int j = 0; j < ntile; ++j, d += 16) {
            const int2 q = *reinterpret_cast<const int2 *>(d);  // row0, live rows (clipped by the dependency warp)
            const bool live = tid < q.y;
            sl_mbar_wait_a(a_full + sb, ph);
            const unsigned int m = sl_lds_u8(a_mask + so);
            double xv[NV][W];
            double acc0 = 0.0, acc1 = 0.0;
            if (__all_sync(0xffffffffu, live && m == FULL)) {
                const long long r8 = 8ll * q.x;
                const char *xb = sl_add_b(xsrc, r8);
                const char *xb2 = NV == 2 ? sl_add_b(xsrc2, r8) : nullptr;
#pragma unroll
                for (int e = 0; e < W; e++) {
                    xv[0][e] = sl_ld_x_off(xb, P.goff8[e]);
                    if (NV == 2) xv[NV - 1][e] = sl_ld_x_off(xb2, P.goff8[e]);
                }
                if (muladd) SlChain<NV, W, 0>::template run<true>(a_val + so, xv, m, false, acc0, acc1);
                else SlChain<NV, W, 0>::template run<false>(a_val + so, xv, m, false, acc0, acc1);
            } else {
                // a dead row's own x entry must stay inside x: clamp its row
                const long long r8 = 8ll * min(q.x, P.n_cols - 1 - tid);
                const char *xb = sl_add_b(xsrc, r8);
                const char *xb2 = NV == 2 ? sl_add_b(xsrc2, r8) : nullptr;
#pragma unroll
                for (int e = 0; e < W; e++) {
                    const long long o = (m & (1u << e)) ? P.goff8[e] : 0ll;  // a slot the row lacks reads its own x entry
                    xv[0][e] = sl_ld_x_off(xb, o);
                    if (NV == 2) xv[NV - 1][e] = sl_ld_x_off(xb2, o);
                }
                if (muladd) SlChain<NV, W, 0>::template run<true>(a_val + so, xv, m, true, acc0, acc1);
                else SlChain<NV, W, 0>::template run<false>(a_val + so, xv, m, true, acc0, acc1);
            }
            if (live) {
                const long long w8 = 8ll * q.x;
                if (last) __stcs(reinterpret_cast<double *>(ydst + w8), acc0);  // nobody in this launch re-reads the last level
                else *reinterpret_cast<double *>(ydst + w8) = acc0;
                if (NV == 2) {
                    if (last) __stcs(reinterpret_cast<double *>(ydst2 + w8), acc1);
                    else *reinterpret_cast<double *>(ydst2 + w8) = acc1;
                }
            }
            __syncwarp();
            if (lane0) sl_mbar_arrive_a(a_empty + sb);  // the stage may be refilled (the warp's reads of it are done)
            so += STAGE;
            sb += 8;
            if (sb == 8u * NS) {
                so = 0;
                sb = 0;
                ph ^= 1u;
            }
        }
        __syncwarp();
        if (lane0) red_release_cta_shared_add(&s_fin[slot], 1u);  // the slot may be reused and the item published
    }
    it_io = it;
    st_io = (int)(sb >> 3);
    ph_io = ph;
}

template <int NV, int W, int NS, int MINB, int EB = 8>
__global__ void __launch_bounds__(SLT_THREADS, MINB) sell_tma_kernel(const SlParams P)
{
    static_assert(W >= 0 && W <= SL_PSLOTS, "pattern width (0: tiles with explicit columns)");
    // pattern tiles: a stage holds the whole blob (mask + W x 256 coefficients); explicit tiles: the lengths + columns
    // prefix of the blob (the coefficients are loaded by the consumers together with the x gathers)
    constexpr int STAGE = W > 0 ? SL_ROWS + 8 * W * SL_ROWS : SL_EXPLICIT_STAGE;
    extern __shared__ __align__(128) unsigned char stages[];  // NS stages
    __shared__ uint64_t s_full[NS];           // the stage's tile has landed (bulk copy complete_tx)
    __shared__ uint64_t s_empty[NS];          // every consumer warp has finished the stage's tile (one arrival per warp)
    __shared__ __align__(16) int s_desc[SL_RING][SL_MAXCHUNK * 16];
    __shared__ int s_hdr[SL_RING][2];
    __shared__ uint64_t s_ready[SL_RING];     // item's inputs complete + descriptors in place
    __shared__ unsigned int s_fin[SL_RING];   // consumer warps that finished the slot's item, counted over the launch
    __shared__ double s_red[SLT_NCW];
    __shared__ SlCta s_cta;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int2 role = __ldg(P.cta_role + blockIdx.x);
    const int level = role.x;
    const int c = role.y;
    const int G = P.team[level];
    const int count = P.count[level];
    const int n_my = c < count ? (count - c + G - 1) / G : 0;
    const int M = P.chunk;
    const int ntl = P.ntl[level];
    if (tid == 0) {
        for (int s = 0; s < SL_RING; s++) {
            mbar_init(&s_ready[s], 1);
            s_fin[s] = 0u;
        }
        for (int s = 0; s < NS; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], SLT_NCW);
        }
        fence_mbar_init();
        s_cta.src[0] = level == 0 ? P.x : P.levels[level - 1];
        s_cta.src[1] = NV == 2 ? (level == 0 ? P.x2 : P.levels2[level - 1]) : nullptr;
        s_cta.dst[0] = P.levels[level];
        s_cta.dst[1] = NV == 2 ? P.levels2[level] : nullptr;
        s_cta.row_end = P.level_rows[level];
        s_cta.last = ((P.flags & 1) && level == P.k - 1) ? 1 : 0;
        s_cta.muladd = P.muladd;
        for (int e = 0; e < 8; e++) s_cta.grel[e] = P.grel[e];
    }
    __syncthreads();

    if (warp == SLT_NCW + 1) {
        // ===== service warp, two duties polled in turn (neither ever blocks the other):
        //  producer  -- tile n of the CTA's stream (items in order, tiles in order) goes to stage n % NS as soon as every
        //               consumer warp has finished the tile NS back;
        //  publisher -- (k > 1) one gpu-scope fence for everything found finished at that moment, then one RED per item.
        //               Consumers bump s_fin[slot] with release.cta after their stores; the acquire here + fence + RED is
        //               cumulative over those stores. =====
        const int4 *tl4p = reinterpret_cast<const int4 *>(P.ltiles[level]);  // quarter 0: blob offset; quarter 1: width, fmt, rp, bytes
        const int *itw = reinterpret_cast<const int *>(P.items[level]);
        int *cnt = P.counters + (size_t)level * P.ngroups;
        const bool last_reader = (P.flags & 1) && level == P.k - 1;
        const uint64_t pol = policy_evict_first();
        int total = 0;  // tiles of the stream: only the level's very last item is short
        if (n_my > 0) {
            const long long t_last = ((long long)c + (long long)(n_my - 1) * G) * M;
            total = (n_my - 1) * M + (int)min((long long)M, (long long)ntl - t_last);
        }
        const int n_pub = P.k > 1 ? n_my : 0;
        long long off_l = 0;  // lane u: blob offset of tile (n & ~31) + u
        int pre_l = STAGE, blob_l = 0;  // ... explicit tiles: bytes of its lengths + columns prefix, of the whole blob
        int grp = 0;          // lane u: group of item (ip & ~31) + u
        int n = 0, ip = 0, off_base = -1, grp_base = -1;
        unsigned long long t_idle = 0;
        while (n < total || ip < n_pub) {
            bool progress = false;
            if (n < total) {
                const int st = n % NS;
                const bool ok = n < NS || mbar_try_wait(&s_empty[st], (uint32_t)(n / NS - 1) & 1u);
                if (ok) {
                    if ((n & ~31) != off_base) {
                        off_base = n & ~31;
                        const int nn = off_base + lane;
                        off_l = 0;
                        if (nn < total) {
                            const long long it = nn / M, j = nn - it * M;
                            const int4 *q = tl4p + (((long long)c + it * G) * M + j) * 4;
                            const int4 q0 = __ldg(q);
                            off_l = ((long long)(unsigned int)q0.x) | ((long long)q0.y << 32);
                            if (W == 0) {
                                const int4 q1 = __ldg(q + 1);
                                pre_l = sl_round_up(2 * q1.z, 128) + 128 * q1.x;
                                blob_l = q1.w;
                            }
                        }
                    }
                    const long long off = __shfl_sync(0xffffffffu, off_l, n & 31);
                    const int pre = W == 0 ? __shfl_sync(0xffffffffu, pre_l, n & 31) : STAGE;
                    const int blob_bytes = W == 0 ? __shfl_sync(0xffffffffu, blob_l, n & 31) : STAGE;
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&s_full[st], (uint32_t)pre);
                        if (last_reader) bulk_g2s_hint(stages + (size_t)st * STAGE, P.blobs + off, (uint32_t)pre, &s_full[st], pol);
                        else bulk_g2s(stages + (size_t)st * STAGE, P.blobs + off, (uint32_t)pre, &s_full[st]);
                        // explicit tiles: the coefficients are read by the consumers; level 0 streams them from HBM, so
                        // pull them into L2 now (NS tiles ahead of their use)
                        if (W == 0 && level == 0 && blob_bytes > pre) sl_prefetch_l2(P.blobs + off + pre, (uint32_t)(blob_bytes - pre));
                    }
                    ++n;
                    progress = true;
                }
            }
            if (ip < n_pub) {
                int nf = 0;  // consecutive finished items (at most to the end of this batch of 32)
                while (ip + nf < n_pub && nf < SL_RING && (nf == 0 || ((ip + nf) & 31) != 0) &&
                       ld_acquire_cta_shared_u32(&s_fin[(ip + nf) % SL_RING]) >= (unsigned int)SLT_NCW * (unsigned int)((ip + nf) / SL_RING + 1))
                    ++nf;
                if (nf > 0) {
                    if ((ip & ~31) != grp_base) {
                        grp_base = ip & ~31;
                        grp = 0;
                        if (grp_base + lane < n_pub) grp = __ldg(itw + ((long long)c + (long long)(grp_base + lane) * G) * 8);
                    }
                    if (lane == 0) __threadfence();
                    __syncwarp();
                    for (int u = 0; u < nf; u++) {
                        const int g = __shfl_sync(0xffffffffu, grp, (ip + u) & 31);
                        if (lane == 0) red_relaxed_gpu_add(cnt + g, 1);
                    }
                    ip += nf;
                    progress = true;
                }
            }
            if (progress) {
                t_idle = 0;
            } else {
                __nanosleep(100);
                const unsigned long long t = sl_now();
                if (t_idle == 0) t_idle = t;
                else if (t - t_idle > SL_TIMEOUT_NS) {
                    if (lane == 0) *P.error = 1;
                    return;
                }
            }
        }
        return;
    }

    if (warp == SLT_NCW) {
        // ===== dependency warp: per item, in order -- descriptors, forward dependencies and back-pressure, ring slot =====
        const int *itw = reinterpret_cast<const int *>(P.items[level]);
        const int4 *tl4 = reinterpret_cast<const int4 *>(P.ltiles[level]);
        const bool fwd = level > 0;
        const bool back = level == 0 && P.bp_level > 0;
        const int *cnt_f = P.counters + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *need_f = P.group_size + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *cnt_b = P.counters + (size_t)(back ? P.bp_level : 0) * P.ngroups;
        const int *need_b = P.group_size + (size_t)(back ? P.bp_level : 0) * P.ngroups;
        int wf = 0, wb = 0;
        bool broken = false;
        // software pipeline over the items: the words and descriptors of item it + 1 are fetched while item it waits
        int iw_n = 0;
        int4 dq_n = make_int4(0, 0, 0, 0);
        auto fetch = [&](int it, int &iw, int4 &dq) {
            const long long i = (long long)c + (long long)it * G;
            iw = 0;
            if (lane < 8) iw = __ldg(itw + i * 8 + lane);
            const long long t0 = i * M;
            const int ntile = (int)min((long long)M, (long long)ntl - t0);
            dq = make_int4(0, 0, 0, 0);
            if (lane < 4 * ntile) dq = __ldg(tl4 + t0 * 4 + lane);
        };
        if (n_my > 0) fetch(0, iw_n, dq_n);
        unsigned long long t_back = 0, t_fwd = 0, t_ring = 0, t_first = 0, t_last = 0;
        for (int it = 0; it < n_my; ++it) {
            const int iw = iw_n;
            const int4 dq = dq_n;
            if (it + 1 < n_my) fetch(it + 1, iw_n, dq_n);
            const long long t0 = ((long long)c + (long long)it * G) * M;
            const int ntile = (int)min((long long)M, (long long)ntl - t0);
            const int s = it % SL_RING;
            const int ghi = __shfl_sync(0xffffffffu, iw, 1);
            const int gback = __shfl_sync(0xffffffffu, iw, 2);
            const unsigned long long ta = P.timing ? sl_now() : 0ull;
            if (!broken) {
                if (back && gback >= wb) wb = sl_wait_groups(cnt_b, need_b, P.epoch, P.ngroups, wb, gback, lane);
            }
            const unsigned long long tb = P.timing ? sl_now() : 0ull;
            if (!broken) {
                if (fwd && ghi >= wf && wb >= 0) wf = sl_wait_groups(cnt_f, need_f, P.epoch, P.ngroups, wf, ghi, lane);
                if (wb < 0 || wf < 0) broken = true;
            }
            const unsigned long long tc = P.timing ? sl_now() : 0ull;
            if (it >= SL_RING && !broken) {
                const unsigned int need = (unsigned int)SLT_NCW * (unsigned int)(it / SL_RING);
                unsigned long long t1 = 0;
                while (ld_acquire_cta_shared_u32(&s_fin[s]) < need) {
                    __nanosleep(100);
                    const unsigned long long t = sl_now();
                    if (t1 == 0) t1 = t;
                    else if (t - t1 > SL_TIMEOUT_NS) { broken = true; break; }
                }
            }
            if (P.timing) {
                const unsigned long long td = sl_now();
                if (it == 0) t_first = ta;
                t_back += tb - ta;
                t_fwd += tc - tb;
                t_ring += td - tc;
                t_last = td;
            }
            if (broken && lane == 0) *P.error = 1;
            if (lane < 4 * SL_MAXCHUNK) {
                int4 dw = dq;
                // quarter 0 of a descriptor = {off lo, off hi, row0, nrows}: rows beyond the level's row prefix are dead
                if ((lane & 3) == 0) dw.w = min(dw.w, s_cta.row_end - dw.z);
                reinterpret_cast<int4 *>(s_desc[s])[lane] = dw;
            }
            // word 6 of a pattern tile's descriptor: 1 = stored with the operator's global pattern (quarter 1 of the descriptor)
            const bool pure = __all_sync(0xffffffffu, !((lane & 3) == 1 && lane < 4 * ntile) || dq.z != 0);
            if (lane == 0) {
                s_hdr[s][0] = ntile;
                s_hdr[s][1] = pure ? 1 : 0;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_ready[s]);  // release.cta: descriptors + everything acquired above
        }
        if (P.timing && lane == 0) {
            long long *o = P.timing + 8 * (size_t)blockIdx.x;
            o[0] = (long long)t_back; o[1] = (long long)t_fwd; o[2] = (long long)t_ring;
            o[3] = (long long)(t_last - t_first); o[4] = n_my; o[5] = level; o[6] = (long long)t_first; o[7] = (long long)t_last;
        }
        return;
    }

    // ===== consumer warps: thread t owns row t of every tile; only the row's x entries are loads in flight =====
    const int t = tid;
    const int cmax = P.n_cols - 1;
    const double *val = reinterpret_cast<const double *>(stages + SL_ROWS) + t;
    // per-CTA constants in registers (the asm loads below clobber memory: the compiler would re-read s_cta per tile)
    const double *const src = s_cta.src[0];
    const double *const src2 = NV == 2 ? s_cta.src[1] : nullptr;
    double *const dst = s_cta.dst[0];
    double *const dst2 = NV == 2 ? s_cta.dst[1] : nullptr;
    const int row_end = s_cta.row_end;
    const bool last = s_cta.last != 0;
    const bool muladd = s_cta.muladd != 0;
    const bool gpat = W > 0 && P.gpat != 0;
    int n_done = 0;
    double dot_acc = 0.0;
    int st = 0;
    uint32_t ph = 0;  // stage of the next tile, parity of its completion
    for (int it = n_done; it < n_my; ++it) {
        if constexpr (W > 0) {
            if (gpat && !P.dot_w) {  // (the fused dot of CG's product stays with the general code below)
                // items of global-pattern tiles; comes back at an item that holds an exception tile (or at the end)
                sl_consume_gpat<NV, W, NS>(P, s_cta, n_my, tid, stages, s_full, s_empty, s_ready, s_fin, s_hdr, s_desc, it, st, ph);
                if (it >= n_my) break;
            }
        }
        const int slot = it % SL_RING;
        mbar_wait(&s_ready[slot], (it / SL_RING) & 1);
        const int ntile = s_hdr[slot][0];
        for (int j = 0; j < ntile; ++j) {
            const int *d = s_desc[slot] + 16 * j;
            const int2 q = *reinterpret_cast<const int2 *>(d + 2);  // row0, nrows
            const int row = q.x + t;
            const bool live = t < q.y && row < row_end;
            double acc0 = 0.0, acc1 = 0.0;
            if constexpr (W > 0) {
                mbar_wait(&s_full[st], ph);
                const unsigned int m = stages[(size_t)st * STAGE + t];
                const double *vs = val + (size_t)st * (STAGE / 8);
                double xv[NV][W];
                {
                    const int4 r0 = *reinterpret_cast<const int4 *>(d + 8), r1 = *reinterpret_cast<const int4 *>(d + 12);
                    const int rel[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    const int rowc = min(row, cmax);
#pragma unroll
                    for (int e = 0; e < W; e++) {
                        const int idx = rowc + ((m & (1u << e)) ? rel[e] : 0);  // a slot the row lacks reads its own x entry
                        xv[0][e] = sl_ld_x(src + idx);
                        if (NV == 2) xv[NV - 1][e] = sl_ld_x(src2 + idx);
                    }
                    if (muladd) {
#pragma unroll
                        for (int e = 0; e < W; e++)
                            if (m & (1u << e)) {
                                const double a = vs[e * SL_ROWS];
                                acc0 = row_op<true>(a, xv[0][e], acc0);
                                if (NV == 2) acc1 = row_op<true>(a, xv[NV - 1][e], acc1);
                            }
                    } else {
#pragma unroll
                        for (int e = 0; e < W; e++)
                            if (m & (1u << e)) {
                                const double a = vs[e * SL_ROWS];
                                acc0 = row_op<false>(a, xv[0][e], acc0);
                                if (NV == 2) acc1 = row_op<false>(a, xv[NV - 1][e], acc1);
                            }
                    }
                }
            } else {
                // explicit columns: d[8 + s] = slots of slice s (this warp's rows: slice `warp`), the slices stored one
                // after the other; lengths and columns from the stage, coefficients and x from global in batches of 8
                const int4 r0 = *reinterpret_cast<const int4 *>(d + 8), r1 = *reinterpret_cast<const int4 *>(d + 12);
                const int rel[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                const int rp = d[6], tot = d[4];
                int soff = 0;
#pragma unroll
                for (int sl = 0; sl < 8; sl++)
                    if (sl < warp) soff += rel[sl];
                const int mine = warp * 32 < rp ? d[8 + warp] : 0;
                const long long off = ((long long)(unsigned int)d[0]) | ((long long)d[1] << 32);
                const int pre = sl_round_up(2 * rp, 128) + 128 * tot;
                const double *vg = reinterpret_cast<const double *>(P.blobs + off + pre) + (size_t)soff * 32 + lane;
                const uint64_t pol = sl_policy(last);
                mbar_wait(&s_full[st], ph);
                const unsigned char *stg = stages + (size_t)st * STAGE;
                const int len = warp * 32 < rp ? (int)reinterpret_cast<const unsigned short *>(stg)[t] : 0;
                const int *cs = reinterpret_cast<const int *>(stg + sl_round_up(2 * rp, 128)) + (size_t)soff * 32 + lane;
                for (int e0 = 0; e0 < mine; e0 += EB) {
                    double a[EB], xv[NV][EB];
                    // every load unconditional (each destination register defined once): a batch that runs past the
                    // slice's last slot re-reads that slot; padding entries carry a valid column (0); neither is used
#pragma unroll
                    for (int u = 0; u < EB; u++) {
                        const int e = min(e0 + u, mine - 1);
                        const int cc = cs[e * 32];
                        a[u] = sl_ld_coef(vg + (size_t)e * 32, pol);
                        xv[0][u] = sl_ld_x(src + cc);
                        if (NV == 2) xv[NV - 1][u] = sl_ld_x(src2 + cc);
                    }
#pragma unroll
                    for (int u = 0; u < EB; u++)
                        if (e0 + u < len) {
                            if (muladd) {
                                acc0 = row_op<true>(a[u], xv[0][u], acc0);
                                if (NV == 2) acc1 = row_op<true>(a[u], xv[NV - 1][u], acc1);
                            } else {
                                acc0 = row_op<false>(a[u], xv[0][u], acc0);
                                if (NV == 2) acc1 = row_op<false>(a[u], xv[NV - 1][u], acc1);
                            }
                        }
                }
            }
            if (live) {
                if (last) __stcs(dst + row, acc0);  // nobody in this launch re-reads the last level
                else dst[row] = acc0;
                if (NV == 2) {
                    if (last) __stcs(dst2 + row, acc1);
                    else dst2[row] = acc1;
                }
                if (NV == 1 && P.dot_w) dot_acc = __fma_rn(P.dot_w[row], acc0, dot_acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);  // the stage may be refilled (the warp's reads of it are done)
            if (++st == NS) {
                st = 0;
                ph ^= 1u;
            }
        }
        __syncwarp();
        if (lane == 0) red_release_cta_shared_add(&s_fin[slot], 1u);  // the slot may be reused and the item published
    }

    if (P.dot_w) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_xor_sync(0xffffffffu, dot_acc, o);
        if (lane == 0) s_red[warp] = dot_acc;
        named_bar_sync(2, SL_ROWS);
        if (warp == 0) {
            __shared__ bool is_last;
            if (lane == 0) {
                double sum = 0.0;
                for (int w = 0; w < SLT_NCW; w++) sum += s_red[w];
                P.partials[blockIdx.x] = sum;
                __threadfence();
                const unsigned int ticket = atomicAdd(P.ticket, 1u);
                is_last = (ticket == gridDim.x - 1);
            }
            __syncwarp();
            if (is_last) {
                __threadfence();
                double sum = 0.0;
                for (int b = lane; b < (int)gridDim.x; b += 32) sum += ld_cg_f64(P.partials + b);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) {
                    *P.dot_out = sum;
                    *P.ticket = 0u;
                }
            }
        }
    }
}

template <int NV, int NS, int MINB>
static sl_fn sl_lookup_tma_w(int w, int *stage_bytes)
{
    *stage_bytes = SL_ROWS + 8 * w * SL_ROWS;
    switch (w) {
    case 1: return sell_tma_kernel<NV, 1, NS, MINB>;
    case 2: return sell_tma_kernel<NV, 2, NS, MINB>;
    case 3: return sell_tma_kernel<NV, 3, NS, MINB>;
    case 4: return sell_tma_kernel<NV, 4, NS, MINB>;
    case 5: return sell_tma_kernel<NV, 5, NS, MINB>;
    case 6: return sell_tma_kernel<NV, 6, NS, MINB>;
    case 7: return sell_tma_kernel<NV, 7, NS, MINB>;
    case 8: return sell_tma_kernel<NV, 8, NS, MINB>;
    }
    return nullptr;
}
// stages per CTA: 3 (four CTAs of 48 registers per SM) or 4 (three CTAs of 64 registers); two right-hand sides: 3 CTAs;
// w = 0: explicit-column tiles (3 stages of 24 KB, three CTAs per SM)
static SlLaunch sl_lookup_tma(int nv, int w, int ns, int *smem)
{
    SlLaunch L;
    L.threads = SLT_THREADS;
    L.launch_regs = 0;  // no register hand-over in this kernel
    if (w == 0) {
        // entries of a row whose loads are issued together (EB): 16 with two CTAs per SM (a 15-entry FEM row costs one
        // memory round trip per level instead of two); option sell_tma = 8: 8 with three CTAs, 17: 16 with three CTAs
        // measured on the RCM-ordered tet-P1 operator (8.1 M rows, k = 4): 16 / two CTAs 0.80 ms, 8 / three CTAs 1.10 ms,
        // 16 / three CTAs (spills) 2.09 ms, four products 0.95 ms (profiles/r02_c4_variants.txt) -> 16 / two is the default
        if (nv == 2) L.fn = sell_tma_kernel<2, 0, 3, 2>;
        else if (ns == 8) L.fn = sell_tma_kernel<1, 0, 3, 3>;
        else if (ns == 17) L.fn = sell_tma_kernel<1, 0, 3, 3, 16>;
        else L.fn = sell_tma_kernel<1, 0, 3, 2, 16>;
        *smem = SL_EXPLICIT_STAGE * 3;
        return L;
    }
    int sb = 0;
    if (nv == 2) { L.fn = sl_lookup_tma_w<2, 3, 3>(w, &sb); ns = 3; }
    else if (ns >= 4) { L.fn = sl_lookup_tma_w<1, 4, 3>(w, &sb); ns = 4; }
    else { L.fn = sl_lookup_tma_w<1, 3, 4>(w, &sb); ns = 3; }
    *smem = sb * ns;
    return L;
}

// -----------------------------------------------------------------------------------------------
// host side: tiling + pattern detection + blobs
// -----------------------------------------------------------------------------------------------
// Uninitialised byte buffer (std::vector would zero a gigabyte on one thread before the threaded packer overwrites it).
struct SlRawBuf {
    std::unique_ptr<unsigned char[]> p;
    size_t n = 0;
    void alloc(size_t bytes) { p.reset(new unsigned char[bytes]); n = bytes; }
    unsigned char *data() { return p.get(); }
    const unsigned char *data() const { return p.get(); }
    size_t size() const { return n; }
};

struct SellHost {
    std::vector<nsk_tile> tiles;   // {row0, nrows, nz0, nz1}: input of nsk_wave_deps
    std::vector<SlTile> stiles;
    SlRawBuf blobs;                // every tile's bytes are zeroed by the thread that fills the tile
    size_t blob_bytes = 0;
    int n_pattern = 0;
    int uniform_width = 0;  // > 0: every tile is a pattern tile stored with this many slots
    int explicit_prefix = 0;  // > 0: every tile has explicit columns; largest lengths + columns prefix of a blob (bytes)
    int global_pattern = 0;   // 1: every tile is stored with ONE column pattern (grel, uniform_width slots)
    int grel[SL_PSLOTS] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static void sl_parallel(int n, const std::function<void(int)> &body)
{
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int t0 = next.fetch_add(64);
            if (t0 >= n) return;
            for (int t = t0; t < std::min(n, t0 + 64); t++) body(t);
        }
    };
    const int nth = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (int i = 1; i < nth; i++) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
}

// Host-only packer.  Returns "" on success, else why the operator is not stored as sliced-ELL tiles.
static std::string sl_pack_host(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                                const std::vector<int> &breaks, SellHost &out)
{
    if (n == 0 || nnz == 0) return "empty operator";
    (void)n_cols;
    std::vector<nsk_tile> &tiles = out.tiles;
    tiles.clear();
    {
        size_t bi = 0;
        int r = 0;
        while (r < n) {
            while (bi < breaks.size() && breaks[bi] <= r) bi++;
            const int seg_end = bi < breaks.size() ? std::min(n, breaks[bi]) : n;
            const int rows = std::min(SL_ROWS, seg_end - r);
            tiles.push_back(nsk_tile{r, rows, ptrow[r], ptrow[r + rows]});
            r += rows;
        }
    }
    const int ntiles = (int)tiles.size();
    std::vector<SlTile> &st = out.stiles;
    st.assign(ntiles, SlTile());
    // 1. per tile: format, geometry, size
    std::atomic<int> too_long(0);
    std::vector<char> unordered((size_t)ntiles, 0);  // tile whose rows do not list their columns in ascending order
    std::vector<std::array<int, 8>> sw_all((size_t)ntiles);  // the slices' widths (explicit layout)
    sl_parallel(ntiles, [&](int t) {
        const nsk_tile &tl = tiles[t];
        SlTile &d = st[t];
        d.row0 = tl.row0;
        d.nrows = tl.nrows;
        d.rp = sl_round_up(tl.nrows, 32);
        int width = 0, sw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < tl.nrows; r++) {
            const int len = ptrow[tl.row0 + r + 1] - ptrow[tl.row0 + r];
            width = std::max(width, len);
            sw[r >> 5] = std::max(sw[r >> 5], len);
        }
        if (width > 65535) { too_long.store(1); return; }
        d.width = width;
        bool pattern = width > 0 && width <= SL_PSLOTS;
        int rel[SL_PSLOTS] = {0, 0, 0, 0, 0, 0, 0, 0};
        int pw = 0;  // slots of the pattern
        if (pattern) {
            // rows with ascending columns (the usual case): the pattern is the sorted union of the rows' offsets -- rows
            // on a domain face lack different neighbours, and no single row need have them all
            bool ascending = true;
            for (int r = 0; r < tl.nrows && pattern && ascending; r++) {
                const int p0 = ptrow[tl.row0 + r], p1 = ptrow[tl.row0 + r + 1];
                for (int j = p0; j < p1 && pattern; j++) {
                    if (j > p0 && indcol[j] <= indcol[j - 1]) { ascending = false; break; }
                    const int o = indcol[j] - (tl.row0 + r);
                    int e = 0;
                    while (e < pw && rel[e] < o) e++;
                    if (e < pw && rel[e] == o) continue;
                    if (pw == SL_PSLOTS) { pattern = false; break; }
                    for (int u = pw; u > e; u--) rel[u] = rel[u - 1];
                    rel[e] = o;
                    pw++;
                }
            }
            if (pattern && !ascending) {
                unordered[t] = 1;
                // any other entry order: the first row of full width is the pattern, every row must be a sub-pattern
                // of it with slots ascending in entry order (the chain's order is the row's storage order)
                int rref = 0;
                while (ptrow[tl.row0 + rref + 1] - ptrow[tl.row0 + rref] != width) rref++;
                const int p = ptrow[tl.row0 + rref];
                pw = width;
                for (int e = 0; e < width; e++) rel[e] = indcol[p + e] - (tl.row0 + rref);
                for (int r = 0; r < tl.nrows && pattern; r++) {
                    const int p0 = ptrow[tl.row0 + r], p1 = ptrow[tl.row0 + r + 1];
                    int e = 0;
                    for (int j = p0; j < p1; j++) {
                        while (e < pw && rel[e] + tl.row0 + r != indcol[j]) e++;
                        if (e == pw) { pattern = false; break; }
                        e++;
                    }
                }
            }
            if (pattern) d.width = width = pw;
        }
        if (width == 0) pattern = true;  // a tile of empty rows: nothing stored but the (zero) masks
        if (pattern) {
            d.fmt = SL_FMT_PATTERN;
            for (int e = 0; e < 8; e++) d.rel[e] = rel[e];
            d.bytes = sl_blob_bytes_pattern(width);
        } else {
            d.fmt = SL_FMT_EXPLICIT;
            int tot = 0;
            for (int s = 0; s < 8; s++) { d.rel[s] = sw[s]; tot += sw[s]; }
            d.width = tot;  // explicit tiles: slots summed over the slices (sizes the column block)
            d.bytes = sl_blob_bytes_explicit(d.rp, tot);
        }
        sw_all[t] = {sw[0], sw[1], sw[2], sw[3], sw[4], sw[5], sw[6], sw[7]};
    });
    if (too_long.load()) return "a row is longer than 65535 entries";
    size_t total = 0;
    int n_pattern = 0, wmax = 0;
    for (int t = 0; t < ntiles; t++) {
        n_pattern += st[t].fmt == SL_FMT_PATTERN;
        if (st[t].fmt == SL_FMT_PATTERN) wmax = std::max(wmax, st[t].width);
    }
    // An operator made of pattern tiles only gets ONE width: narrower tiles (lines on a domain face lack a neighbour
    // line) are padded with slots no row has, so the streaming kernel runs one straight-line instance over all of them.
    out.uniform_width = 0;
    out.explicit_prefix = 0;
    out.global_pattern = 0;
    if (n_pattern == ntiles && wmax >= 1) {
        // ONE pattern for (nearly) the whole operator: the union of the tiles' patterns when it still has at most 8 slots
        // (any stencil on a box: the tiles on a domain face only lack some of the interior tiles' offsets); otherwise the
        // union of the most frequent patterns that fits (a distributed slab: the tiles next to a ghost ring reference it
        // at offsets of their own and stay EXCEPTIONS with their own pattern).  The kernel takes the global offsets from
        // its parameters (constant bank); a warp of a global tile whose 32 rows have every slot runs straight-line code
        // without masks.  Global tiles need rows with ascending columns (slot order = entry order); rp = 1 marks a global tile.
        {
            std::map<std::array<int, 9>, int> freq;  // {pw, rel[0..7]} -> tiles
            for (int t = 0; t < ntiles; t++) {
                if (unordered[t]) continue;  // (slots in entry order, not by offset: always an exception)
                std::array<int, 9> key = {st[t].width, 0, 0, 0, 0, 0, 0, 0, 0};
                for (int e = 0; e < st[t].width; e++) key[1 + e] = st[t].rel[e];
                freq[key]++;
            }
            std::vector<std::pair<int, std::array<int, 9>>> order;
            for (auto &kv : freq) order.push_back({kv.second, kv.first});
            std::stable_sort(order.begin(), order.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
            std::vector<int> uni;
            for (auto &o : order) {
                std::vector<int> u2 = uni;
                for (int e = 0; e < o.second[0]; e++) {
                    auto it = std::lower_bound(u2.begin(), u2.end(), o.second[1 + e]);
                    if (it == u2.end() || *it != o.second[1 + e]) u2.insert(it, o.second[1 + e]);
                }
                if ((int)u2.size() <= wmax) uni.swap(u2);  // never wider than the widest tile: a slot costs every tile 2 KB
            }
            if ((int)uni.size() == wmax) {
                // (the straight-line path runs all W slots of the kernel instance: the global pattern must fill them)
                out.global_pattern = 1;
                wmax = (int)uni.size();
                for (int e = 0; e < SL_PSLOTS; e++) out.grel[e] = e < wmax ? uni[e] : 0;
                for (int t = 0; t < ntiles; t++) {
                    bool sub = !unordered[t];
                    for (int e = 0; e < st[t].width && sub; e++) sub = std::binary_search(uni.begin(), uni.end(), st[t].rel[e]);
                    st[t].rp = sub ? 1 : 0;
                    if (sub)
                        for (int e = 0; e < SL_PSLOTS; e++) st[t].rel[e] = out.grel[e];
                }
            }
        }
        out.uniform_width = wmax;
        for (int t = 0; t < ntiles; t++) {
            st[t].width = wmax;  // rel[] beyond the tile's own slots is already zero
            st[t].bytes = sl_blob_bytes_pattern(wmax);
        }
    } else {
        // Anything else is stored with explicit columns THROUGHOUT (a pattern tile here and there would only make the
        // kernels branch): the staged kernel copies a tile's lengths + columns (a contiguous prefix of its blob).
        n_pattern = 0;
        for (int t = 0; t < ntiles; t++) {
            SlTile &d = st[t];
            int tot = 0;
            for (int s = 0; s < 8; s++) { d.rel[s] = sw_all[t][s]; tot += d.rel[s]; }
            d.fmt = SL_FMT_EXPLICIT;
            d.width = tot;
            d.bytes = sl_blob_bytes_explicit(d.rp, tot);
            out.explicit_prefix = std::max(out.explicit_prefix, sl_round_up(2 * d.rp, 128) + 128 * tot);
        }
    }
    for (int t = 0; t < ntiles; t++) {
        st[t].off = (long long)total;
        total += (size_t)st[t].bytes;
    }
    // slice padding: refuse operators whose rows are so ragged that the tiles would outweigh CSR by half
    if ((double)total > 1.5 * (12.0 * (double)nnz + 4.0 * n) + 65536.0) return "row lengths too ragged for sliced-ELL tiles";
    out.blob_bytes = total;
    out.n_pattern = n_pattern;
    out.blobs.alloc(total + 128);
    memset(out.blobs.data() + total, 0, 128);
    // 2. the blobs
    sl_parallel(ntiles, [&](int t) {
        const nsk_tile &tl = tiles[t];
        const SlTile &d = st[t];
        unsigned char *b = out.blobs.data() + d.off;
        memset(b, 0, (size_t)d.bytes);  // padding slots / rows read as zeros
        if (d.fmt == SL_FMT_PATTERN) {
            unsigned char *mask = b;
            double *val = reinterpret_cast<double *>(b + SL_ROWS);
            for (int r = 0; r < tl.nrows; r++) {
                const int p0 = ptrow[tl.row0 + r], p1 = ptrow[tl.row0 + r + 1];
                unsigned int m = 0;
                int e = 0;
                for (int j = p0; j < p1; j++) {
                    while (d.rel[e] + tl.row0 + r != indcol[j]) e++;  // phase 1 proved that a slot matches
                    m |= 1u << e;
                    val[(size_t)e * SL_ROWS + r] = coef[j];
                    e++;
                }
                mask[r] = (unsigned char)m;
            }
        } else {
            int tot = 0, soff[8];
            for (int s = 0; s < 8; s++) { soff[s] = tot; tot += d.rel[s]; }
            unsigned short *len = reinterpret_cast<unsigned short *>(b);
            int *col = reinterpret_cast<int *>(b + sl_round_up(2 * d.rp, 128));
            double *val = reinterpret_cast<double *>(b + sl_round_up(2 * d.rp, 128) + sl_round_up(128 * tot, 128));
            for (int r = 0; r < tl.nrows; r++) {
                const int p0 = ptrow[tl.row0 + r], p1 = ptrow[tl.row0 + r + 1];
                len[r] = (unsigned short)(p1 - p0);
                const size_t base = (size_t)soff[r >> 5] * 32 + (size_t)(r & 31);
                for (int j = p0; j < p1; j++) {
                    col[base + (size_t)(j - p0) * 32] = indcol[j];
                    val[base + (size_t)(j - p0) * 32] = coef[j];
                }
            }
        }
    });
    return "";
}

// Expands tile t back to CSR rows (host; tests and the GPU packer's cross-check).
static void sl_expand_tile(const SlTile &d, const unsigned char *blobs, std::vector<int> &len_out, std::vector<int> &col_out,
                           std::vector<double> &val_out)
{
    const unsigned char *b = blobs + d.off;
    len_out.assign(d.nrows, 0);
    col_out.clear();
    val_out.clear();
    if (d.fmt == SL_FMT_PATTERN) {
        const unsigned char *mask = b;
        const double *val = reinterpret_cast<const double *>(b + SL_ROWS);
        for (int r = 0; r < d.nrows; r++) {
            for (int e = 0; e < d.width; e++)
                if ((mask[r] >> e) & 1) {
                    col_out.push_back(d.row0 + r + d.rel[e]);
                    val_out.push_back(val[(size_t)e * SL_ROWS + r]);
                    len_out[r]++;
                }
        }
    } else {
        int tot = 0, soff[8];
        for (int s = 0; s < 8; s++) { soff[s] = tot; tot += d.rel[s]; }
        const unsigned short *len = reinterpret_cast<const unsigned short *>(b);
        const int *col = reinterpret_cast<const int *>(b + sl_round_up(2 * d.rp, 128));
        const double *val = reinterpret_cast<const double *>(b + sl_round_up(2 * d.rp, 128) + sl_round_up(128 * tot, 128));
        for (int r = 0; r < d.nrows; r++) {
            const size_t base = (size_t)soff[r >> 5] * 32 + (size_t)(r & 31);
            len_out[r] = len[r];
            for (int e = 0; e < (int)len[r]; e++) {
                col_out.push_back(col[base + (size_t)e * 32]);
                val_out.push_back(val[base + (size_t)e * 32]);
            }
        }
    }
}

// -----------------------------------------------------------------------------------------------
// level schedule (host only): per-level tile lists, items, groups, dependencies, CTA roles
// -----------------------------------------------------------------------------------------------
struct SlSchedule {
    int k = 0, chunk = 1, gi = 1, ngroups = 0;
    std::vector<std::vector<int>> ltile;   // per level: tile ids in position order (inside the level's row prefix)
    std::vector<std::vector<SlItem>> items;
    std::vector<int> gsize;                // [k][ngroups]
    std::vector<int> teams;
    std::vector<int2> roles;
    bool same_lists = true;                // every level has level 0's tile list
};

// tile_at_pos: tile id at each position of global row order; pmax[t]: last position tile t's columns refer to.
// lead: how far (positions) level 0 may run ahead of level k-1 per hop.  teams[] in: requested team sizes.
static void sl_build_schedule(const std::vector<SlTile> &tiles, const std::vector<int> &tile_at_pos, const std::vector<int> &pmax,
                              const std::vector<int> &lr, int k, int chunk, int lead, int interleave, std::vector<int> teams,
                              SlSchedule &S)
{
    const int ntiles = (int)tiles.size();
    S.k = k;
    S.chunk = chunk;
    S.gi = std::max(1, WF_GROUP / chunk);
    S.ltile.assign(k, std::vector<int>());
    S.items.assign(k, std::vector<SlItem>());
    std::vector<std::vector<int>> lpos(k);  // positions of the level's tiles (ascending)
    S.same_lists = true;
    for (int l = 0; l < k; l++) {
        for (int pos = 0; pos < ntiles; pos++) {
            const int t = tile_at_pos[pos];
            if (tiles[t].row0 >= lr[l]) continue;  // outside this level's row prefix (distributed shrink)
            S.ltile[l].push_back(t);
            lpos[l].push_back(pos);
        }
        if (l > 0 && S.ltile[l] != S.ltile[0]) S.same_lists = false;
    }
    int maxitems = 0;
    for (int l = 0; l < k; l++) maxitems = std::max(maxitems, ((int)S.ltile[l].size() + chunk - 1) / chunk);
    S.ngroups = std::max(1, (maxitems + S.gi - 1) / S.gi);
    S.gsize.assign((size_t)k * S.ngroups, 0);
    const int hold = k > 1 ? (k - 1) * lead : -1;
    for (int l = 0; l < k; l++) {
        const int nt = (int)S.ltile[l].size();
        const int ni = (nt + chunk - 1) / chunk;
        S.items[l].resize(ni);
        for (int i = 0; i < ni; i++) {
            SlItem &it = S.items[l][i];
            it.group = i / S.gi;
            it.ghi = -1;
            it.gback = -1;
            it.pad[0] = it.pad[1] = 0;
            const int a = i * chunk, b = std::min(nt, a + chunk);
            if (l > 0) {
                int pm = -1;
                for (int u = a; u < b; u++) pm = std::max(pm, pmax[S.ltile[l][u]]);
                // tiles of the level below at positions <= pm
                const int j = (int)(std::upper_bound(lpos[l - 1].begin(), lpos[l - 1].end(), pm) - lpos[l - 1].begin()) - 1;
                if (j >= 0) it.ghi = (j / chunk) / S.gi;
            }
            if (l == 0 && hold >= 0) {
                const int p = lpos[l][a] - hold;
                if (p > 0) {
                    // items of level k-1 that lie entirely below position p
                    const int j = (int)(std::lower_bound(lpos[k - 1].begin(), lpos[k - 1].end(), p) - lpos[k - 1].begin());
                    it.gback = (j / chunk) / S.gi - 1;
                }
            }
            // blobs of consecutive tile ids are contiguous
            bool contiguous = true;
            for (int u = a + 1; u < b; u++) contiguous = contiguous && S.ltile[l][u] == S.ltile[l][u - 1] + 1;
            it.pf_off = tiles[S.ltile[l][a]].off;
            it.pf_bytes = 0;
            if (contiguous) {
                const SlTile &lastt = tiles[S.ltile[l][b - 1]];
                it.pf_bytes = (int)(lastt.off + lastt.bytes - it.pf_off);
            }
            S.gsize[(size_t)l * S.ngroups + it.group]++;
        }
    }
    for (int l = 0; l < k; l++) teams[l] = std::max(1, std::min(teams[l], std::max(1, (int)S.items[l].size())));
    S.teams = teams;
    S.roles.clear();
    int total = 0;
    for (int l = 0; l < k; l++) total += teams[l];
    std::vector<int> given(k, 0);
    if (interleave) {
        for (int b = 0; b < total; b++) {
            int best = -1;
            double bestv = 0.0;
            for (int l = 0; l < k; l++) {  // the level furthest behind its share
                if (given[l] >= teams[l]) continue;
                const double v = (double)(b + 1) * teams[l] / total - given[l];
                if (best < 0 || v > bestv) { best = l; bestv = v; }
            }
            S.roles.push_back(make_int2(best, given[best]++));
        }
    } else {
        for (int l = 0; l < k; l++)
            for (int i = 0; i < teams[l]; i++) S.roles.push_back(make_int2(l, i));
    }
}

// CPU model of the kernel's protocol (tests): every CTA opens its items strictly in order (the dependency warp is
// head-of-line blocking), an item opens when groups <= ghi of the level below and groups <= gback of level k-1 have
// reached their sizes, up to `ring` items per CTA are open at once and finish in random order, but are PUBLISHED in
// order.  Returns the number of items published; the schedule is sound iff that equals the total.  reads[t] = tiles
// whose rows tile t's nonzeros reference; *violations counts items opened while a tile they read was unpublished at
// the level below (and inside that level's prefix).
static long long sl_simulate(const SlSchedule &S, int ring, unsigned seed, const std::vector<std::vector<int>> *reads,
                             const std::vector<SlTile> *tiles, const std::vector<int> *lr, long long *violations)
{
    const int k = S.k, grid = (int)S.roles.size();
    const int ntiles_all = reads ? (int)reads->size() : 0;
    std::vector<std::vector<char>> tile_done(reads ? k : 0, std::vector<char>((size_t)ntiles_all, 0));
    std::vector<std::vector<int>> cnt(k, std::vector<int>(S.ngroups, 0));
    std::vector<int> water(k, 0);
    auto advance = [&](int l) {
        while (water[l] < S.ngroups && cnt[l][water[l]] >= S.gsize[(size_t)l * S.ngroups + water[l]]) water[l]++;
    };
    for (int l = 0; l < k; l++) advance(l);
    struct Open { int idx; bool finished; };
    std::vector<int> next(grid, 0);
    std::vector<std::vector<Open>> open(grid);  // in opening order
    long long done = 0, total = 0, bad = 0;
    for (int l = 0; l < k; l++) total += (long long)S.items[l].size();
    unsigned rng = seed * 2654435761u + 12345u;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    std::vector<int> order(grid);
    for (int b = 0; b < grid; b++) order[b] = b;
    bool progress = true, force = false;
    while (done < total && (progress || !force)) {
        force = !progress;
        progress = false;
        for (int i = grid - 1; i > 0; i--) std::swap(order[i], order[rnd() % (unsigned)(i + 1)]);
        for (int oi = 0; oi < grid; oi++) {
            const int b = order[oi];
            const int level = S.roles[b].x, c = S.roles[b].y, G = S.teams[level];
            // finish one open item (random pick), then publish the finished prefix
            if (!open[b].empty() && (force || (rnd() & 1) || (int)open[b].size() >= ring)) {
                std::vector<int> cand;
                for (int u = 0; u < (int)open[b].size(); u++)
                    if (!open[b][u].finished) cand.push_back(u);
                if (!cand.empty()) {
                    open[b][cand[rnd() % (unsigned)cand.size()]].finished = true;
                    progress = true;
                }
                while (!open[b].empty() && open[b].front().finished) {
                    const int idx = open[b].front().idx;
                    const SlItem &it = S.items[level][idx];
                    open[b].erase(open[b].begin());
                    cnt[level][it.group]++;
                    if (reads) {
                        const int a = idx * S.chunk, e = std::min((int)S.ltile[level].size(), a + S.chunk);
                        for (int u = a; u < e; u++) tile_done[level][(size_t)S.ltile[level][u]] = 1;
                    }
                    advance(level);
                    done++;
                    progress = true;
                }
            }
            const long long idx = (long long)c + (long long)next[b] * G;
            if ((int)open[b].size() < ring && idx < (long long)S.items[level].size()) {
                const SlItem &it = S.items[level][(size_t)idx];
                const bool fwd_ok = level == 0 || it.ghi < 0 || water[level - 1] > it.ghi || water[level - 1] >= S.ngroups;
                const bool back_ok = it.gback < 0 || water[k - 1] > it.gback || water[k - 1] >= S.ngroups;
                if (fwd_ok && back_ok) {
                    if (reads && level > 0) {
                        const int a = (int)idx * S.chunk, e = std::min((int)S.ltile[level].size(), a + S.chunk);
                        for (int u = a; u < e; u++)
                            for (int d : (*reads)[(size_t)S.ltile[level][u]])
                                if ((*tiles)[(size_t)d].row0 < (*lr)[level - 1] && !tile_done[level - 1][(size_t)d]) bad++;
                    }
                    open[b].push_back(Open{(int)idx, false});
                    next[b]++;
                    progress = true;
                }
            }
        }
    }
    if (violations) *violations = bad;
    return done;
}

static std::vector<int> sl_team_sizes(int resident, int k, int w0_pct)
{
    std::vector<int> teams(k, 0);
    const double wsum = (double)w0_pct + 100.0 * (k - 1);
    int used = 0;
    for (int l = 1; l < k; l++) { teams[l] = std::max(1, (int)(resident * 100.0 / wsum)); used += teams[l]; }
    teams[0] = std::max(1, resident - used);
    return teams;
}

// -----------------------------------------------------------------------------------------------
// host-only access for the CPU test-suite: pack, expand, simulate
// -----------------------------------------------------------------------------------------------
struct nsk_sell_host_s {
    SellHost H;
    std::string why;
    int n = 0, n_cols = 0;
    std::vector<int> ptrow, indcol;
};

NSK_API int nsk_sell_host_create(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                                 void **out)
{
    if (!out || !ptrow || (nnz > 0 && (!indcol || !coef))) return NSK_ERR_INVALID;
    nsk_sell_host_s *h = new nsk_sell_host_s();
    h->n = n;
    h->n_cols = n_cols;
    h->ptrow.assign(ptrow, ptrow + n + 1);
    h->indcol.assign(indcol, indcol + nnz);
    h->why = sl_pack_host(n, n_cols, nnz, ptrow, indcol, coef, std::vector<int>(), h->H);
    *out = h;
    return NSK_OK;
}

NSK_API const char *nsk_sell_host_why(void *handle) { return static_cast<nsk_sell_host_s *>(handle)->why.c_str(); }

NSK_API int nsk_sell_host_stats(void *handle, int64_t *bytes, int64_t *ntiles, int64_t *pattern_tiles)
{
    nsk_sell_host_s *h = static_cast<nsk_sell_host_s *>(handle);
    if (!h->why.empty()) return NSK_ERR_UNSUPPORTED;
    if (bytes) *bytes = (int64_t)h->H.blob_bytes;
    if (ntiles) *ntiles = (int64_t)h->H.stiles.size();
    if (pattern_tiles) *pattern_tiles = h->H.n_pattern;
    return NSK_OK;
}

// Global pattern of the operator: returns its number of slots (0: none), the offsets in rel[8] and the number of tiles
// stored with it (the others are exceptions with a pattern of their own).
NSK_API int nsk_sell_host_global_pattern(void *handle, int *rel, int64_t *global_tiles)
{
    nsk_sell_host_s *h = static_cast<nsk_sell_host_s *>(handle);
    if (!h->why.empty() || !h->H.global_pattern) return 0;
    int64_t g = 0;
    for (const SlTile &d : h->H.stiles) g += d.rp != 0;
    if (global_tiles) *global_tiles = g;
    for (int e = 0; e < SL_PSLOTS && rel; e++) rel[e] = h->H.grel[e];
    return h->H.uniform_width;
}

// Expands the tiles back to CSR (row pointers, global columns, values).  Arrays sized n+1 / nnz by the caller.
NSK_API int nsk_sell_host_expand(void *handle, int *ptrow, int *indcol, double *coef)
{
    nsk_sell_host_s *h = static_cast<nsk_sell_host_s *>(handle);
    if (!h->why.empty()) return NSK_ERR_UNSUPPORTED;
    int64_t k = 0;
    ptrow[0] = 0;
    std::vector<int> len, col;
    std::vector<double> val;
    for (const SlTile &d : h->H.stiles) {
        sl_expand_tile(d, h->H.blobs.data(), len, col, val);
        size_t q = 0;
        for (int r = 0; r < d.nrows; r++) {
            for (int e = 0; e < len[r]; e++, q++) {
                indcol[k] = col[q];
                coef[k] = val[q];
                k++;
            }
            ptrow[d.row0 + r + 1] = (int)k;
        }
    }
    return NSK_OK;
}

// Builds the level schedule exactly like the GPU path (exact dependencies from the columns, natural row order,
// optional per-level row prefixes) and runs the CPU protocol model.  Returns the number of items that did NOT get
// published (0 = sound), -1000000 - v when v items opened before their inputs were published, or a negative status.
NSK_API long long nsk_sell_host_simulate(void *handle, int k, int chunk, int lead_slack_tiles, int resident, int w0_pct,
                                         int interleave, int ring, const int *level_rows, unsigned seed, long long *items_out,
                                         int *reach_out, int pmax_bias)
{
    nsk_sell_host_s *h = static_cast<nsk_sell_host_s *>(handle);
    if (!h->why.empty()) return NSK_ERR_UNSUPPORTED;
    if (k < 1 || k > NSK_MAX_K || resident < k || ring < 1 || chunk < 1 || chunk > SL_MAXCHUNK) return NSK_ERR_INVALID;
    const std::vector<SlTile> &tiles = h->H.stiles;
    const int ntiles = (int)tiles.size();
    std::vector<int> row0s(ntiles), tile_at_pos(ntiles), pmax(ntiles, 0);
    for (int t = 0; t < ntiles; t++) { row0s[t] = tiles[t].row0; tile_at_pos[t] = t; }
    std::vector<std::vector<int>> reads((size_t)ntiles);
    int reach = 0;
    for (int t = 0; t < ntiles; t++) {
        std::vector<int> &rd = reads[(size_t)t];
        int mx = -1;
        for (int j = h->ptrow[tiles[t].row0]; j < h->ptrow[tiles[t].row0 + tiles[t].nrows]; j++) {
            const int g = h->indcol[j];
            if (g >= h->n) continue;  // ghost entry of x: read by level 0 only
            const int d = (int)(std::upper_bound(row0s.begin(), row0s.end(), g) - row0s.begin()) - 1;
            mx = std::max(mx, d);
            if (rd.empty() || rd.back() != d) rd.push_back(d);
        }
        std::sort(rd.begin(), rd.end());
        rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
        pmax[t] = mx < 0 ? t : mx;
        reach = std::max(reach, pmax[t] - t);
        pmax[t] = std::max(-1, pmax[t] + pmax_bias);  // tests weaken the dependencies on purpose to see the model object
    }
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : h->n;
    // a negative slack undercuts the safe minimum on purpose: the test-suite checks that the model then reports a deadlock
    const int lead = std::max(1, reach + 1 + WF_GROUP + chunk + lead_slack_tiles);
    SlSchedule S;
    sl_build_schedule(tiles, tile_at_pos, pmax, lr, k, chunk, lead, interleave, sl_team_sizes(resident, k, w0_pct), S);
    long long violations = 0, total = 0;
    for (int l = 0; l < k; l++) total += (long long)S.items[l].size();
    const long long done = sl_simulate(S, ring, seed, &reads, &tiles, &lr, &violations);
    if (items_out) *items_out = total;
    if (reach_out) *reach_out = reach;
    if (violations > 0) return -1000000 - violations;
    return total - done;
}

NSK_API void nsk_sell_host_destroy(void *handle) { delete static_cast<nsk_sell_host_s *>(handle); }

// -----------------------------------------------------------------------------------------------
// device side: the operator's tiles (cached per operator), level plans, launches
// -----------------------------------------------------------------------------------------------
struct SlPlan {
    int k = 0, resident = 0, chunk = 0, w0_pct = 0, interleave = 0, l2_pct = 0, lead_pct = 0, nv = 1;
    std::vector<int> level_rows;
    bool rejected = false;
    int grid = 0, ngroups = 0, reach = 0, lead = 0, epoch = 0;
    std::vector<int> teams, count, ntl;
    std::vector<const SlItem *> d_items;    // per level (into d_item_buf)
    std::vector<const SlTile *> d_ltiles;   // per level (into d_ltile_buf, or the operator's own array)
    SlItem *d_item_buf = nullptr;
    SlTile *d_ltile_buf = nullptr;
    int *d_counters = nullptr, *d_group_size = nullptr;
    int2 *d_roles = nullptr;
};

struct SellOp {
    bool ok = false;
    std::string why;
    int ntiles = 0, n_pattern = 0, uniform_width = 0, explicit_prefix = 0, global_pattern = 0;
    int grel[SL_PSLOTS] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t blob_bytes = 0;
    unsigned char *d_blobs = nullptr;
    SlTile *d_tiles = nullptr;
    std::vector<SlTile> h_tiles;
    nsk_tiling csr_view;
    std::vector<SlPlan> plans;
    int *h_error = nullptr, *d_error = nullptr;  // host-mapped watchdog flag
};

static std::map<nsk_csr_t, SellOp *> g_sell;
static std::mutex g_sell_mu;

void nsk_sell_free(nsk_csr_t A)
{
    SellOp *op = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_sell_mu);
        auto it = g_sell.find(A);
        if (it == g_sell.end()) return;
        op = it->second;
        g_sell.erase(it);
    }
    for (SlPlan &p : op->plans) {
        if (p.d_item_buf) cudaFree(p.d_item_buf);
        if (p.d_ltile_buf) cudaFree(p.d_ltile_buf);
        if (p.d_counters) cudaFree(p.d_counters);
        if (p.d_group_size) cudaFree(p.d_group_size);
        if (p.d_roles) cudaFree(p.d_roles);
    }
    if (op->d_blobs) cudaFree(op->d_blobs);
    if (op->d_tiles) cudaFree(op->d_tiles);
    if (op->h_error) cudaFreeHost(op->h_error);
    delete op;
}

static SellOp *sl_get(nsk_csr_t A)
{
    {
        std::lock_guard<std::mutex> lk(g_sell_mu);
        auto it = g_sell.find(A);
        if (it != g_sell.end()) return it->second;
    }
    SellOp *op = new SellOp();
    {
        std::lock_guard<std::mutex> lk(g_sell_mu);
        g_sell[A] = op;
    }
    const int n = A->n;
    const std::vector<int> &ptrow = nsk_csr_host_ptrow(A);
    if (n == 0 || A->nnz == 0) { op->why = "empty operator"; return op; }
    if ((int)ptrow.size() != n + 1) { op->why = "host row pointers missing"; return op; }
    // the caller's host arrays are gone: read the entries back once and pack on the host
    // (uninitialised buffers: a std::vector would zero 1.4 GB for a 256^3 operator before the copy overwrites it)
    std::unique_ptr<int[]> indcol(new int[(size_t)A->nnz]);
    std::unique_ptr<double[]> coef(new double[(size_t)A->nnz]);
    if (cudaMemcpy(indcol.get(), A->d_indcol, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(coef.get(), A->d_coef, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost) != cudaSuccess) {
        op->why = "reading the operator back failed";
        return op;
    }
    SellHost H;
    op->why = sl_pack_host(n, A->n_cols, A->nnz, ptrow.data(), indcol.get(), coef.get(), A->breaks, H);
    indcol.reset();
    coef.reset();
    if (!op->why.empty()) return op;
    const int ntiles = (int)H.tiles.size();
    if (cudaMalloc(&op->d_blobs, H.blobs.size()) != cudaSuccess ||
        cudaMalloc(&op->d_tiles, sizeof(SlTile) * (size_t)ntiles) != cudaSuccess ||
        cudaHostAlloc(&op->h_error, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&op->d_error, op->h_error, 0) != cudaSuccess) {
        op->why = "allocation of the sliced-ELL operator failed";
        cudaGetLastError();
        return op;
    }
    *op->h_error = 0;
    cudaMemcpy(op->d_blobs, H.blobs.data(), H.blobs.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(op->d_tiles, H.stiles.data(), sizeof(SlTile) * (size_t)ntiles, cudaMemcpyHostToDevice);
    op->ntiles = ntiles;
    op->n_pattern = H.n_pattern;
    op->uniform_width = H.uniform_width;
    op->explicit_prefix = H.explicit_prefix;
    op->global_pattern = H.global_pattern;
    for (int e = 0; e < SL_PSLOTS; e++) op->grel[e] = H.grel[e];
    op->blob_bytes = H.blob_bytes;
    op->h_tiles.swap(H.stiles);
    op->csr_view.tile_rows = SL_ROWS;
    op->csr_view.ntiles = ntiles;
    op->csr_view.nlong = 0;
    op->csr_view.h_tiles.swap(H.tiles);
    op->ok = true;
    return op;
}

// tiles per item.  Generic kernel: the tiles a consumer thread takes at once (3 for one right-hand side, 2 for two).
// Streaming kernel: only the granularity of dependencies and publication.
static int sl_plan_chunk(nsk_ctx_t ctx, int nv, bool stream)
{
    int chunk = (int)ctx->opt.sell_chunk;
    const int cap = stream ? SL_MAXCHUNK : sl_max_chunk(nv);
    if (chunk <= 0) chunk = stream ? 4 : cap;
    return std::max(1, std::min(chunk, cap));
}

static SlPlan *sl_plan(nsk_csr_t A, SellOp *op, int k, const int *level_rows, int resident, int nv, bool stream, const char **why)
{
    nsk_ctx_t ctx = A->ctx;
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : A->n;
    const int chunk = sl_plan_chunk(ctx, nv, stream);
    const int w0_pct = k > 1 ? (ctx->opt.pipe_w0_pct > 0 ? (int)ctx->opt.pipe_w0_pct : 100) : 100;
    const int interleave = ctx->opt.pipe_interleave ? 1 : 0;
    const int l2_pct = (int)ctx->opt.wave_l2_pct;
    const int lead_pct = (int)ctx->opt.wave_slack_pct;
    for (SlPlan &p : op->plans)
        if (p.k == k && p.resident == resident && p.level_rows == lr && p.chunk == chunk && p.w0_pct == w0_pct &&
            p.interleave == interleave && p.l2_pct == l2_pct && p.lead_pct == lead_pct && p.nv == nv) {
            if (p.rejected) { *why = "level window exceeds the L2 budget"; return nullptr; }
            return &p;
        }
    const int ntiles = op->ntiles;
    SlPlan p;
    p.k = k; p.resident = resident; p.level_rows = lr; p.chunk = chunk; p.w0_pct = w0_pct; p.interleave = interleave;
    p.l2_pct = l2_pct; p.lead_pct = lead_pct; p.nv = nv;
    std::vector<int> tile_at_pos(ntiles), pmax(ntiles, 0);
    for (int t = 0; t < ntiles; t++) tile_at_pos[t] = t;
    if (k > 1) {
        WaveDeps D;
        if (!nsk_wave_deps(A, op->csr_view, D, why)) return nullptr;
        tile_at_pos = D.tile_at_pos;
        for (int t = 0; t < ntiles; t++) pmax[t] = std::min(ntiles - 1, (D.ghi[t] + 1) * WF_GROUP - 1);
        p.reach = D.reach;
        // lead = how far level 0 may run ahead of level k-1, per hop: the pattern's reach + one completion group + one
        // chunk at least, plus slack for the publish -> poll latency and the items in flight.  The window that must
        // stay L2-resident is (k-1) * lead tiles of blobs plus the level vectors over it; by default the slack is
        // whatever the L2 budget allows.
        const double tile_bytes = (double)op->blob_bytes / ntiles + 8.0 * SL_ROWS * (k + 1) * nv;
        // budget: share of L2 the window may take.  Pattern operators, one right-hand side: 92 % (256^3, k = 4, final kernel:
        // 0.5130 ms at 86 %, 0.5087 at 88, 0.5050 at 90, 0.5035 at 92, 0.5028 at 94 -- profiles/r02_sweep_final.txt; a slab
        // with ghost rows: 535.8 us at 88, 535.6 at 92, 541.4 at 96, 549.7 at 100); two right-hand sides and explicit-column
        // operators (whose gathers compete for L2): 88 %
        const double l2_default = (op->uniform_width > 0 && nv == 1) ? 92.0 : 88.0;
        const double budget = (l2_pct > 0 ? (double)l2_pct : l2_default) / 100.0 * (double)ctx->prop.l2CacheSize;
        const int lead_min = D.reach + 1 + WF_GROUP + chunk;
        if (lead_pct >= 0)
            p.lead = lead_min + (int)((double)lead_pct / 100.0 * 2.0 * (resident / k) * chunk + 0.999);
        else
            p.lead = std::max(lead_min + 2 * WF_GROUP, (int)(budget / ((double)(k - 1) * tile_bytes)));
        const double window = (double)(k - 1) * p.lead * tile_bytes;
        // the items a team has in flight (team x chunk tiles) must fit between two levels besides the reach, or every
        // level waits on the loop latency: below that the caller fuses fewer levels per launch
        const int min_slack = (resident / k) * chunk;
        const bool thin = lead_pct < 0 && p.lead < ntiles && p.lead - lead_min < min_slack;
        if (window > budget * 1.0001 || thin) {
            *why = "level window exceeds the L2 budget";
            p.rejected = true;
            op->plans.push_back(p);
            return nullptr;
        }
    }
    SlSchedule S;
    sl_build_schedule(op->h_tiles, tile_at_pos, pmax, lr, k, chunk, p.lead, interleave, sl_team_sizes(resident, k, w0_pct), S);
    p.teams = S.teams;
    p.ngroups = S.ngroups;
    p.grid = (int)S.roles.size();
    p.count.resize(k);
    p.ntl.resize(k);
    size_t nitems = 0, nlt = 0;
    for (int l = 0; l < k; l++) {
        p.count[l] = (int)S.items[l].size();
        p.ntl[l] = (int)S.ltile[l].size();
        nitems += S.items[l].size();
        nlt += S.ltile[l].size();
    }
    bool identity = S.same_lists && p.ntl[0] == ntiles;
    for (int t = 0; t < ntiles && identity; t++) identity = S.ltile[0][t] == t;
    if (cudaMalloc(&p.d_roles, sizeof(int2) * S.roles.size()) != cudaSuccess ||
        cudaMalloc(&p.d_item_buf, sizeof(SlItem) * (nitems + 1)) != cudaSuccess ||
        cudaMalloc(&p.d_counters, sizeof(int) * (size_t)k * S.ngroups) != cudaSuccess ||
        cudaMalloc(&p.d_group_size, sizeof(int) * (size_t)k * S.ngroups) != cudaSuccess ||
        (!identity && cudaMalloc(&p.d_ltile_buf, sizeof(SlTile) * (nlt + 1)) != cudaSuccess)) {
        *why = "plan allocation failed";
        cudaGetLastError();
        return nullptr;
    }
    p.d_items.resize(k);
    p.d_ltiles.resize(k);
    size_t io = 0, lo = 0;
    std::vector<SlTile> lt;
    for (int l = 0; l < k; l++) {
        p.d_items[l] = p.d_item_buf + io;
        if (!S.items[l].empty())
            cudaMemcpy(p.d_item_buf + io, S.items[l].data(), sizeof(SlItem) * S.items[l].size(), cudaMemcpyHostToDevice);
        io += S.items[l].size();
        if (identity) {
            p.d_ltiles[l] = op->d_tiles;
        } else {
            lt.resize(S.ltile[l].size());
            for (size_t u = 0; u < lt.size(); u++) lt[u] = op->h_tiles[(size_t)S.ltile[l][u]];
            p.d_ltiles[l] = p.d_ltile_buf + lo;
            if (!lt.empty()) cudaMemcpy(p.d_ltile_buf + lo, lt.data(), sizeof(SlTile) * lt.size(), cudaMemcpyHostToDevice);
            lo += lt.size();
        }
    }
    cudaMemset(p.d_counters, 0, sizeof(int) * (size_t)k * S.ngroups);
    cudaMemcpy(p.d_group_size, S.gsize.data(), sizeof(int) * S.gsize.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_roles, S.roles.data(), sizeof(int2) * S.roles.size(), cudaMemcpyHostToDevice);
    op->plans.push_back(p);
    return &op->plans.back();
}

static int sl_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2, double *const *d_levels2,
                  nsk_mode mode, const int *level_rows, const double *dot_w, int dot_slot)
{
    nsk_ctx_t ctx = A->ctx;
    const int nv = d_x2 ? 2 : 1;
    SellOp *op = sl_get(A);
    if (!op->ok) { nsk_set_error(ctx, "sliced-ELL path not applicable: %s", op->why.c_str()); return NSK_ERR_UNSUPPORTED; }
    if (*op->h_error) {
        nsk_set_error(ctx, "fused matrix-powers kernel: a bounded wait expired in an earlier launch (results invalid)");
        return NSK_ERR_CUDA;
    }
    // all-pattern operators: streaming consumers (option sell_stream: 0 = default depth 3, n = depth n, < 0 = off);
    // everything else: an item at a time
    const int depth = ctx->opt.sell_stream == 0 ? 3 : (int)ctx->opt.sell_stream;
    const bool stream = op->uniform_width > 0 && depth >= 2;
    const int rows = ctx->opt.sell_rows == 2 ? 2 : 1;
    // all-pattern operators, default: coefficients staged by bulk copies (option sell_tma: 0 = default 4 stages, n = n
    // stages, < 0 = off -> the register kernels above)
    // operators with explicit columns whose lengths + columns fit a stage: the same kernel, columns staged
    const bool tma_e = op->uniform_width == 0 && op->explicit_prefix > 0 && op->explicit_prefix <= SL_EXPLICIT_STAGE;
    const bool tma = (op->uniform_width > 0 || tma_e) && ctx->opt.sell_tma >= 0;
    int smem = 0;
    const SlLaunch L = tma ? sl_lookup_tma(nv, op->uniform_width, ctx->opt.sell_tma == 0 ? 3 : (int)ctx->opt.sell_tma, &smem)
                     : stream ? sl_lookup_stream(nv, op->uniform_width, depth, rows)
                              : sl_lookup(nv, sl_plan_chunk(ctx, nv, false));
    sl_fn fn = L.fn;
    if (tma) NSK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (!tma) {
        // the register hand-over between the warpgroups only adds up when the kernel got the register count its launch
        // bounds imply (ptxas pins it to that when setmaxnreg is used): refuse to launch otherwise rather than hang
        cudaFuncAttributes fa;
        NSK_CUDA(ctx, cudaFuncGetAttributes(&fa, fn));
        if (fa.numRegs != L.launch_regs) {
            nsk_set_error(ctx, "sliced-ELL kernel was built with %d registers per thread, expected %d", fa.numRegs, L.launch_regs);
            return NSK_ERR_UNSUPPORTED;
        }
    }
    int per_sm = 0;
    NSK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, L.threads, smem));
    if (ctx->opt.sell_ctas_per_sm > 0) per_sm = std::min(per_sm, (int)ctx->opt.sell_ctas_per_sm);
    int resident = ctx->prop.multiProcessorCount * per_sm;
    if (ctx->opt.sell_max_ctas > 0) resident = std::max(k, std::min(resident, (int)ctx->opt.sell_max_ctas));  // tests: long streams per CTA
    if (resident < k) { nsk_set_error(ctx, "sliced-ELL path: fewer resident CTAs than levels"); return NSK_ERR_UNSUPPORTED; }
    const char *why = "";
    SlPlan *plan = sl_plan(A, op, k, level_rows, resident, nv, stream || tma, &why);
    if (!plan) { nsk_set_error(ctx, "fused matrix powers not applicable: %s", why); return NSK_ERR_UNSUPPORTED; }
    int maxcount = 0;
    for (int l = 0; l < k; l++) maxcount = std::max(maxcount, plan->count[l]);
    if (maxcount == 0) return NSK_OK;
    if (dot_w) NSK_REQUIRE(ctx, k == 1 && plan->grid <= NSK_MAX_PARTIALS, "fused dot: k = 1 and a bounded grid");
    if (plan->epoch >= 100000000) {  // counters are monotone over launches: start over long before they overflow
        NSK_CUDA(ctx, cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (size_t)k * plan->ngroups, ctx->stream));
        plan->epoch = 0;
    }
    plan->epoch++;
    SlParams P;
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < k; l++) {
        P.items[l] = plan->d_items[l];
        P.ltiles[l] = plan->d_ltiles[l];
        P.count[l] = plan->count[l];
        P.ntl[l] = plan->ntl[l];
        P.team[l] = plan->teams[l];
        P.level_rows[l] = plan->level_rows[l];
        P.levels[l] = d_levels[l];
        P.levels2[l] = d_levels2 ? d_levels2[l] : nullptr;
    }
    P.x = d_x;
    P.x2 = d_x2;
    P.blobs = op->d_blobs;
    P.counters = plan->d_counters;
    P.group_size = plan->d_group_size;
    P.cta_role = plan->d_roles;
    P.ngroups = plan->ngroups;
    P.n_cols = A->n_cols;
    P.k = k;
    P.chunk = plan->chunk;
    P.epoch = plan->epoch;
    P.bp_level = k > 1 ? k - 1 : -1;
    P.muladd = mode == NSK_EXACT_MULADD ? 1 : 0;
    P.flags = ctx->opt.sell_flags >= 0 ? (int)ctx->opt.sell_flags : 3;
    P.pf_dist = ctx->opt.sell_pf_dist > 0 ? (int)ctx->opt.sell_pf_dist : 2;
    P.gpat = (op->global_pattern && ctx->opt.sell_geom != 2) ? 1 : 0;  // option sell_geom = 2: masked path everywhere
    for (int e = 0; e < SL_PSLOTS; e++) {
        P.grel[e] = op->grel[e];
        P.goff8[e] = 8ll * op->grel[e];
    }
    P.error = op->d_error;
    long long *d_timing = nullptr;
    if (ctx->opt.pk_timing && tma) {
        NSK_CUDA(ctx, cudaMalloc(&d_timing, sizeof(long long) * 8 * (size_t)plan->grid));
        NSK_CUDA(ctx, cudaMemset(d_timing, 0, sizeof(long long) * 8 * (size_t)plan->grid));
    }
    P.timing = d_timing;
    P.dot_w = dot_w;
    P.partials = ctx->d_partials;
    P.ticket = ctx->d_ticket;
    P.dot_out = dot_w ? ctx->d_scalars + dot_slot : nullptr;
    if (k > 1) {
        // CTAs of different levels wait on each other: co-residency must be guaranteed, not assumed
        void *args[] = {&P};
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)fn, dim3(plan->grid), dim3(L.threads), args, (size_t)smem, ctx->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
            cudaGetLastError();
            plan->epoch--;
            nsk_set_error(ctx, "fused matrix powers: the grid cannot be made co-resident on this device now");
            return NSK_ERR_UNSUPPORTED;
        }
        NSK_CUDA(ctx, e);
    } else {
        fn<<<plan->grid, L.threads, smem, ctx->stream>>>(P);
        NSK_CUDA(ctx, cudaGetLastError());
    }
    ctx->launches++;
    if (d_timing) {
        std::vector<long long> h((size_t)plan->grid * 8);
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        NSK_CUDA(ctx, cudaMemcpy(h.data(), d_timing, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(d_timing);
        long long t0 = -1, t1 = 0;
        for (int b = 0; b < plan->grid; b++)
            if (h[8 * b + 4] > 0) { t0 = t0 < 0 ? h[8 * b + 6] : std::min(t0, h[8 * b + 6]); t1 = std::max(t1, h[8 * b + 7]); }
        for (int l = 0; l < k; l++) {
            double sb = 0, sf = 0, sr = 0, st = 0, si = 0, first = 1e30, lastt = 0;
            int nb = 0;
            for (int b = 0; b < plan->grid; b++)
                if (h[8 * b + 5] == l && h[8 * b + 4] > 0) {
                    sb += h[8 * b]; sf += h[8 * b + 1]; sr += h[8 * b + 2]; st += h[8 * b + 3]; si += h[8 * b + 4]; nb++;
                    first = std::min(first, (double)(h[8 * b + 6] - t0)); lastt = std::max(lastt, (double)(h[8 * b + 7] - t0));
                }
            if (nb)
                fprintf(stderr, "sell_timing k=%d level %d: %d CTAs, %.1f items each, per item ns: back-pressure wait %.0f | forward wait %.0f | "
                                "ring (consumers) wait %.0f | cycle %.0f ; level active %.1f .. %.1f us of %.1f\n", k, l, nb, si / nb, sb / si,
                        sf / si, sr / si, st / si, first / 1e3, lastt / 1e3, (t1 - t0) / 1e3);
        }
    }
    ctx->last_sell[0] = op->uniform_width;
    ctx->last_sell[1] = plan->reach;
    ctx->last_sell[2] = plan->lead;
    ctx->last_sell[3] = plan->grid;
    ctx->last_sell[4] = plan->d_ltile_buf ? 0 : 1;  // 1: every level walks the operator's own tile array
    ctx->last_sell[5] = op->ntiles;
    ctx->last_sell[6] = tma ? 1 : 0;
    ctx->last_sell[7] = plan->ngroups;
    return NSK_OK;
}

int nsk_sell_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode, const int *level_rows,
                 const double *dot_w, int dot_slot)
{
    return sl_run(A, k, d_x, d_levels, nullptr, nullptr, mode, level_rows, dot_w, dot_slot);
}

int nsk_sell_run2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2, double *const *d_levels2,
                  nsk_mode mode, const int *level_rows)
{
    return sl_run(A, k, d_x, d_levels, d_x2, d_levels2, mode, level_rows, nullptr, -1);
}

bool nsk_sell_applicable(nsk_csr_t A)
{
    if (A->n == 0 || A->nnz == 0) return false;
    return sl_get(A)->ok;
}

// Explicit-column operator whose tiles' lengths + columns fit a stage of the staged kernel (unstructured FEM operators with
// up to ~23 entries per row on average).
bool nsk_sell_explicit_staged(nsk_csr_t A)
{
    if (A->n == 0 || A->nnz == 0) return false;
    SellOp *op = sl_get(A);
    return op->ok && op->uniform_width == 0 && op->explicit_prefix > 0 && op->explicit_prefix <= SL_EXPLICIT_STAGE;
}

// Every tile is a pattern tile of one width: the staged-coefficient kernel applies (stencils, regular bands).
bool nsk_sell_uniform(nsk_csr_t A)
{
    if (A->n == 0 || A->nnz == 0) return false;
    SellOp *op = sl_get(A);
    return op->ok && op->uniform_width > 0;
}

// 1 when the watchdog flag of the operator's fused kernel is set (a bounded wait expired); clears it.
int nsk_sell_check_error(nsk_csr_t A)
{
    std::lock_guard<std::mutex> lk(g_sell_mu);
    auto it = g_sell.find(A);
    if (it == g_sell.end() || !it->second->h_error) return 0;
    const int e = *it->second->h_error;
    *it->second->h_error = 0;
    return e;
}

size_t nsk_sell_bytes(nsk_csr_t A)
{
    SellOp *op = sl_get(A);
    return op->ok ? op->blob_bytes + sizeof(SlTile) * (size_t)op->ntiles : 0;
}
