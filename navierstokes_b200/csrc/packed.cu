// packed.cu -- tile-packed operator format + ONE persistent kernel for y = A x and for the fused matrix powers
// A x, A^2 x, ..., A^k x, in which the consumer warps never touch global memory for their inputs.
//
// Replaces SpMV_CSR / _OPT / _FMA / _AVX2 (reference mpk/SpMV.cpp:6-85) and the fused SpM2V_CSR* / SpM3V /
// SpM4V (mpk/SpM2V.cpp:80-332, mpk/SpMVmulti0.cpp:132-221) for operators whose tiles reference x in a few
// contiguous runs (stencils on structured grids, banded / block-banded matrices).  Everything else keeps the
// CSR kernels (spmv_kernels.cu, mpk_pipeline.cu); both are GPU paths.
//
// Why.  Every CSR kernel here gathers x[col] with one global load per nonzero.  ncu on the 256^3 7-point
// operator (profiles/r01_ncu_mpk_wavefront_c_summary.txt) shows those kernels bound by the LATENCY of that
// gather: ~1 us per round trip, one row per thread in flight, the rows in flight per SM capped by the shared
// memory their matrix slices occupy -- 67 k rows/us whatever the schedule, i.e. 0.25 ms per product with HBM at
// 85 % and L2 at 40 %.  The fix is to take the gather off the SM's load path altogether:
//
//   * at plan time every tile (<= T_ROWS consecutive rows) gets the list of contiguous column runs (SEGMENTS)
//     its nonzeros reference -- 3 runs for a 7-point stencil -- and its column indices are rewritten as 16-bit
//     offsets into the concatenation of those runs;
//   * the tile is stored as one contiguous BLOB in slot-major (sliced-ELL) order: lens[r], lcol[e][r], val[e][r];
//     per-row nonzero order is untouched, so the fma chain is bit-identical to the reference's row loop
//     (10 B/nnz instead of CSR's 12: the operator costs 17 % less HBM traffic as a side effect);
//   * a stage = blob + x runs, all moved by 1-D bulk async copies (TMA): one for the blob, one per run.  The
//     copies of the x runs are issued by the dependency warp the moment the tile's inputs are complete, so the
//     latency of reading another SM's output is hidden by the stage ring, not by resident warps;
//   * consumer threads (one row each) read lens/lcol/val/x from shared memory only: conflict-free for stencils
//     (consecutive rows -> consecutive addresses in every array), no long-scoreboard stall at all.
//
// Matrix powers = level pipeline: the resident CTAs are split into k teams (role table), team l computes power l+1
// only, tile by tile in global row order.  A tile of level l starts when the level l-1 tile groups covering its
// column range are complete (per-group completion counters; the dependency warp polls them BEFORE it waits for its
// stage, so the round trip overlaps the consumers' work, then fences the async proxy and issues the x copies);
// finished tiles are published by a separate warp (one gpu-scope fence for everything finished at that moment);
// level 0 may not run more than a WINDOW ahead of level k-1, and the window -- sized from the L2 budget -- is what
// keeps the k-1 re-reads of every blob in L2 (ncu: 1.93 GB of HBM traffic for k = 4 on 256^3, the operator is read
// once).  The last reader loads blobs evict-first and the last level is stored with a streaming hint.  k = 1 is the
// plain product (no counters, no fences), optionally with the fused dot of CG.  NV = 2 instances carry a second
// right-hand side through the same stages (nsk_mpk_multi: s-step bases of p and r in one sweep).
// The host half (tiling, runs, blobs, level schedule) is plain C++ and is tested without a GPU, including a CPU model
// of this protocol (tests/test_packed_host.py).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"
#include "stream_common.cuh"
#include "wave_common.h"

using namespace nskptx;

std::vector<int> &nsk_csr_host_ptrow(nsk_csr_t A);

constexpr int PK_MAXSEG = 8;

struct PkTile {                   // 96 bytes; words 0..23 are fetched one per lane by the dependency warp
    long long blob_off;           // w0,1   bytes from the blob base, 16-byte aligned
    int blob_bytes;               // w2     multiple of 16
    int nseg;                     // w3
    int row0, nrows, xlen, tail;  // w4..7  xlen: total doubles of all runs (each run a multiple of 2); tail: 1 + the x
                                  //        buffer slot of column n_cols-1 when n_cols is odd and the tile needs it
                                  //        (bulk copies move 16-byte granules: that one element is copied by hand)
    int seg_start[PK_MAXSEG];     // w8..15 first column of the run (even)
    int seg_lenoff[PK_MAXSEG];    // w16..23 length | offset in the stage's x buffer << 16 (doubles)
};
static_assert(sizeof(PkTile) == 96, "PkTile is 24 words");

struct PkItem {            // 32 bytes = two 16-byte loads; one per (level, tile)
    int tile, pos;         // tile index; position in global row order (completion group = pos / WF_GROUP)
    int ghi, gback;        // forward: groups [0, ghi] of level l-1 complete; back-pressure: [0, gback] of level l+1
    long long blob_off;
    int blob_bytes, pad;
};
static_assert(sizeof(PkItem) == 32, "PkItem is two int4");

struct PkParams {
    const PkItem *items[NSK_MAX_K];  // per level, ascending position
    int count[NSK_MAX_K];
    const PkTile *tiles;
    const unsigned char *blobs;
    int *counters;          // [k][ngroups], zeroed before the launch (unused for k = 1)
    const int *group_size;  // [k][ngroups]
    int ngroups;
    int n_cols;             // length of x / of the level vectors
    const double *x;
    double *levels[NSK_MAX_K];
    const double *x2;            // second right-hand side (NV = 2 kernels): same operator, same schedule, its own
    double *levels2[NSK_MAX_K];  // powers -- the blob is read once for both
    int level_rows[NSK_MAX_K];
    int k;
    int team[NSK_MAX_K];   // CTAs of each level (level 0 streams from HBM and gets more stages in flight)
    const int2 *cta_role;  // [grid] {level, index within the level's team}
    int flags;       // experiment switches: 1 = evict-first / streaming hints for data nobody re-reads, 2 = poll without
                     // sleeping, 4 = publish with red.release.gpu instead of fence + relaxed red
    int bp_global;   // 1: only level 0 is held back, by level k-1 (one window for the whole pipeline); 0: level l by l+1
    // optional stage-cycle instrumentation (tools/pk_timing.py): 8 sums of nanoseconds + item count per CTA
    unsigned long long *timing;
    // fused dot <dot_w, levels[0]> (k = 1 only; CG: p.Ap)
    const double *dot_w;
    double *partials;
    unsigned int *ticket;
    double *dot_out;
};

// blob header (16 ints at the start of every blob)
enum { PKH_ROW0 = 0, PKH_NROWS, PKH_WIDTH, PKH_RP, PKH_OFF_LENS, PKH_OFF_LCOL, PKH_OFF_VAL, PKH_FORMAT, PKH_OFF_BASE,
       PKH_WORDS = 16 };
// PKH_FORMAT is always 0 (explicit local columns, lcol u16[width][rp], lens[r] = row length); the pattern-compressed
// format of round 1 moved to sell.cu, where no per-entry index is stored at all.

__host__ __device__ constexpr int pk_round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int pk_blob_bytes(int nrows, int width)
{
    const int rp = pk_round_up(nrows, 32);
    return PKH_WORDS * 4 + 2 * rp + 2 * width * rp + 8 * width * rp;  // every term is a multiple of 16
}

// host-side decoding of a blob: length of row r, local column / value of its entry e
struct PkBlobView {
    const int *hdr;
    const unsigned short *lens, *lcol;
    const double *val;
    int rp;
    explicit PkBlobView(const unsigned char *b)
    {
        hdr = reinterpret_cast<const int *>(b);
        rp = hdr[PKH_RP];
        lens = reinterpret_cast<const unsigned short *>(b + hdr[PKH_OFF_LENS]);
        lcol = reinterpret_cast<const unsigned short *>(b + hdr[PKH_OFF_LCOL]);
        val = reinterpret_cast<const double *>(b + hdr[PKH_OFF_VAL]);
    }
    int len(int r) const { return (int)lens[r]; }
    int col(int e, int r) const { return (int)lcol[(size_t)e * rp + r]; }
    double value(int e, int r) const { return val[(size_t)e * rp + r]; }
};

__device__ __forceinline__ unsigned long long pk_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int pk_wait_groups(const int *cnt, const int *need, int ngroups, int w, int upto, int lane,
                                              bool nosleep = false)
{
    uint32_t spins = 0;
    while (w <= upto) {
        const int g = w + lane;
        bool ok = true;
        if (g < ngroups) ok = ld_acquire_gpu(cnt + g) >= __ldg(need + g);
        const unsigned int bad = __ballot_sync(0xffffffffu, !ok);
        const int adv = bad ? __ffs(bad) - 1 : 32;
        w = min(w + adv, ngroups);
        if (w > upto || w >= ngroups) break;
        if (adv == 0) {
            if (!nosleep) __nanosleep(32);
            if (++spins > (1u << 26)) __trap();  // protocol bug: fail the launch, never hang
        }
    }
    return w;
}

template <int T_ROWS, int BLOB_CAP, int XCAP, int STAGES, int NCW, int MINB, int RPT, int NV, bool MULADD>
__global__ void __launch_bounds__((NCW + 3) * 32, MINB) packed_kernel(const PkParams P)
{
    static_assert(NV == 1 || NV == 2, "one or two right-hand sides");
    constexpr int STAGE_BYTES = BLOB_CAP + NV * XCAP * 8;
    static_assert(BLOB_CAP % 128 == 0 && (XCAP * 8) % 128 == 0, "stage parts keep 128-byte alignment");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGE_BYTES * STAGES);  // blob + x runs landed
    uint64_t *done = full + STAGES;                                                      // consumers stored their rows
    double *red = reinterpret_cast<double *>(done + STAGES);
    unsigned long long *ts = reinterpret_cast<unsigned long long *>(red + 64);  // [STAGES][4] timestamps (timing only)
    // per stage: consumer warps that finished it, counted over the whole launch (monotone, so the publisher can
    // fall behind the stage ring without a phase ever aliasing)
    unsigned int *fin = reinterpret_cast<unsigned int *>(ts + STAGES * 4);
    const bool timing = P.timing != nullptr;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 2);  // producer (blob bytes) + dependency warp (x bytes)
            mbar_init(&done[s], NCW);
            fin[s] = 0u;
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int2 role = __ldg(P.cta_role + blockIdx.x);
    const int level = role.x;
    const int c = role.y;
    const int G = P.team[level];
    const int count = P.count[level];
    const int n_my = c < count ? (count - c + G - 1) / G : 0;
    const int4 *my = reinterpret_cast<const int4 *>(P.items[level]);  // item i of this CTA = my[2 * (c + i * G) ...]

    if (warp == NCW) {
        // ===== producer (one lane).  The blob of a tile depends on no flag: a stage is refilled the moment its
        // previous item is done. =====
        if (lane != 0) return;
        int it_load = 0;
        int4 db = make_int4(0, 0, 0, 0), nb = db;
        if (n_my > 0) db = my[2 * (size_t)c + 1];
        if (n_my > 1) nb = my[2 * (size_t)(c + G) + 1];
        unsigned long long acc[5] = {0, 0, 0, 0, 0};
        // the last level is the last reader of a blob: tell L2 so (the dead copy otherwise ages out of the window's way)
        const bool last_reader = (P.flags & 1) && P.k > 1 && level == P.k - 1;
        const uint64_t pol = policy_evict_first();
        for (; it_load < n_my; ++it_load) {
            const int s = it_load % STAGES;
            if (it_load >= STAGES) {
                mbar_wait(&done[s], ((it_load / STAGES) - 1) & 1);
                if (timing) {
                    const unsigned long long t4 = pk_now();
                    const volatile unsigned long long *v = ts + s * 4;
                    acc[0] += v[1] - v[0];  // blob issue -> x issue (stage waits for its inputs / the dependency warp)
                    acc[1] += v[2] - v[1];  // x issue -> consumers see the stage full (bulk-copy latency)
                    acc[2] += v[3] - v[2];  // consume (warp 0)
                    acc[3] += t4 - v[3];    // warp 0 finished -> producer sees every warp done
                    acc[4] += t4 - v[0];    // whole stage cycle
                }
            }
            const long long off = ((long long)(unsigned int)db.x) | ((long long)db.y << 32);
            mbar_arrive_expect_tx(&full[s], (uint32_t)db.z);
            if (last_reader) bulk_g2s_hint(smem + (size_t)s * STAGE_BYTES, P.blobs + off, (uint32_t)db.z, &full[s], pol);
            else bulk_g2s(smem + (size_t)s * STAGE_BYTES, P.blobs + off, (uint32_t)db.z, &full[s]);
            if (timing) ts[s * 4 + 0] = pk_now();
            db = nb;
            if (it_load + 2 < n_my) nb = my[2 * ((size_t)c + (size_t)(it_load + 2) * G) + 1];
        }
        if (timing) {
            for (int j = 0; j < 5; j++) P.timing[(size_t)blockIdx.x * 16 + j] = acc[j];
            P.timing[(size_t)blockIdx.x * 16 + 8] = (unsigned long long)(n_my > STAGES ? n_my - STAGES : 0);
            P.timing[(size_t)blockIdx.x * 16 + 9] = (unsigned long long)level;
        }
        return;
    }

    if (warp == NCW + 2) {
        // ===== publisher (k > 1): makes finished items visible to the other SMs.  The gpu-scope fence (~0.5 us) is
        // paid here -- off the consumers' path and off the refill path -- ONCE for all items found finished at
        // that moment: consumers bump fin[s] with release.cta after their stores; this warp's acquire of fin[s]
        // followed by fence + RED is cumulative over those stores. =====
        if (P.k <= 1) return;
        int *cnt = P.counters + (size_t)level * P.ngroups;
        int pos_cur = 0;  // lane u: position of item it0 + u
        unsigned long long tf = 0;
        for (int it = 0; it < n_my;) {
            const int j = it & 31;
            if (j == 0) {
                pos_cur = 0;
                if (it + lane < n_my) pos_cur = my[2 * ((size_t)c + (size_t)(it + lane) * G)].y;
            }
            int n = 0;
            uint32_t spins = 0;
            for (;;) {  // items finish in order per stage; take every consecutive finished one (at most to the batch end)
                const int i2 = it + n;
                bool ok = false;
                if (i2 < n_my && (n == 0 || (i2 & 31) != 0) && n < STAGES) {
                    const unsigned int need = (unsigned int)NCW * (unsigned int)(i2 / STAGES + 1);
                    ok = ld_acquire_cta_shared_u32(&fin[i2 % STAGES]) >= need;
                }
                ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
                if (ok) { ++n; continue; }
                if (n > 0) break;
                __nanosleep(32);
                if (++spins > (1u << 26)) __trap();
            }
            const unsigned long long t5 = timing ? pk_now() : 0ull;
            const bool rel = (P.flags & 4) != 0;
            if (lane == 0 && !rel) __threadfence();
            __syncwarp();
            for (int u = 0; u < n; u++) {
                const int pos = __shfl_sync(0xffffffffu, pos_cur, (it + u) & 31);
                if (lane == 0) {
                    if (rel && u == 0) red_release_gpu_add(cnt + pos / WF_GROUP, 1);
                    else red_relaxed_gpu_add(cnt + pos / WF_GROUP, 1);
                }
            }
            if (timing) tf += pk_now() - t5;
            it += n;
        }
        if (timing && lane == 0) P.timing[(size_t)blockIdx.x * 16 + 5] = tf;
        return;
    }

    if (warp == NCW + 1) {
        // ===== dependency warp: once item `it`'s forward inputs (level - 1) are complete and the next level is
        // not more than `lead` tiles behind, it copies the tile's x runs into the stage -- the consumers only wait
        // for bytes. =====
        const bool fwd = level > 0;
        const bool back = P.bp_global ? (level == 0 && P.k > 1) : (level < P.k - 1);
        const int lb = P.bp_global ? P.k - 1 : level + 1;  // the level whose progress holds this one back
        const int *cnt_f = P.counters + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *need_f = P.group_size + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *cnt_b = P.counters + (size_t)(back ? lb : 0) * P.ngroups;
        const int *need_b = P.group_size + (size_t)(back ? lb : 0) * P.ngroups;
        const double *src = level == 0 ? P.x : P.levels[level - 1];
        const double *src2 = NV == 2 ? (level == 0 ? P.x2 : P.levels2[level - 1]) : nullptr;
        const int *tw = reinterpret_cast<const int *>(P.tiles);
        int wf = 0, wb = 0;
        unsigned long long w_done = 0, w_dep = 0;
        int4 cur = make_int4(0, 0, 0, -1), nxt = cur;  // lane u: item it0 + u (cur) and it0 + 32 + u (nxt)
        if (lane < n_my) cur = my[2 * ((size_t)c + (size_t)lane * G)];
        if (32 + lane < n_my) nxt = my[2 * ((size_t)c + (size_t)(32 + lane) * G)];
        // lane u < 24: word u of the PkTile of item it (cw), it + 1 (nw), it + 2 (fw): fetched TWO items ahead, so
        // the descriptor's L2 (first sweep: HBM) latency never sits between a stage becoming free and its x copies
        int cw = 0, nw = 0, fw = 0;
        auto tile_of = [&](int i, int j0) {  // tile of item i, given that `cur` holds items j0 .. j0 + 31
            const int a = __shfl_sync(0xffffffffu, cur.x, (i - j0) & 31);
            const int b = __shfl_sync(0xffffffffu, nxt.x, (i - j0) & 31);
            return i - j0 < 32 ? a : b;
        };
        if (n_my > 0) {
            const int t0 = tile_of(0, 0);
            if (lane < 24) cw = __ldg(tw + (size_t)t0 * 24 + lane);
        }
        if (n_my > 1) {
            const int t1 = tile_of(1, 0);
            if (lane < 24) nw = __ldg(tw + (size_t)t1 * 24 + lane);
        }
        for (int it = 0; it < n_my; ++it) {
            const int j = it & 31;
            if (j == 0 && it > 0) {
                cur = nxt;
                nxt = make_int4(0, 0, 0, -1);
                if (it + 32 + lane < n_my) nxt = my[2 * ((size_t)c + (size_t)(it + 32 + lane) * G)];
            }
            if (it + 2 < n_my) {
                const int tn = tile_of(it + 2, it - j);
                if (lane < 24) fw = __ldg(tw + (size_t)tn * 24 + lane);
            }
            const int ghi = __shfl_sync(0xffffffffu, cur.z, j);
            const int gback = __shfl_sync(0xffffffffu, cur.w, j);
            const int s = it % STAGES;
            // inputs first: the poll (an L2 round trip whenever the watermark has to move) overlaps the consumers'
            // work on the item that still occupies this stage
            const unsigned long long t0 = timing ? pk_now() : 0ull;
            if (back && gback >= wb) wb = pk_wait_groups(cnt_b, need_b, P.ngroups, wb, gback, lane, (P.flags & 2) != 0);
            if (fwd && ghi >= wf) wf = pk_wait_groups(cnt_f, need_f, P.ngroups, wf, ghi, lane, (P.flags & 2) != 0);
            const unsigned long long t1 = timing ? pk_now() : 0ull;
            if (it >= STAGES) mbar_wait(&done[s], ((it / STAGES) - 1) & 1);  // the stage's x buffer is free
            if (timing) { w_dep += t1 - t0; w_done += pk_now() - t1; }
            if (fwd) fence_proxy_async_global();  // acquired generic-proxy writes -> visible to the bulk copies below
            const int nseg = __shfl_sync(0xffffffffu, cw, 3);
            const int xlen = __shfl_sync(0xffffffffu, cw, 6);
            const int start = __shfl_sync(0xffffffffu, cw, 8 + (lane & 7));
            const int lenoff = __shfl_sync(0xffffffffu, cw, 16 + (lane & 7));
            const int tail = __shfl_sync(0xffffffffu, cw, 7);
            if (lane == 0) {
                if (tail) {  // odd vector length: the last element cannot ride a 16-byte bulk copy
                    double *xs = reinterpret_cast<double *>(smem + (size_t)s * STAGE_BYTES + BLOB_CAP);
                    xs[tail - 1] = ld_cg_f64(src + P.n_cols - 1);
                    if (NV == 2) xs[XCAP + tail - 1] = ld_cg_f64(src2 + P.n_cols - 1);
                }
                mbar_arrive_expect_tx(&full[s], (uint32_t)xlen * 8u * NV);  // release: orders the stores above too
            }
            __syncwarp();
            if ((lane & 7) < nseg && (lane >> 3) < NV) {  // lanes 0..7: runs of the first vector, 8..15: of the second
                const int len = lenoff & 0xffff, xoff = (lenoff >> 16) & 0xffff;
                const int v = lane >> 3;
                bulk_g2s(smem + (size_t)s * STAGE_BYTES + BLOB_CAP + ((size_t)v * XCAP + (size_t)xoff) * 8,
                         (v == 0 ? src : src2) + start, (uint32_t)len * 8u, &full[s]);
            }
            if (timing && lane == 0) ts[s * 4 + 1] = pk_now();
            cw = nw;
            nw = fw;
        }
        if (timing && lane == 0) {
            P.timing[(size_t)blockIdx.x * 16 + 6] = w_done;  // dependency warp: waiting for its stage to be free
            P.timing[(size_t)blockIdx.x * 16 + 7] = w_dep;   // ... and for the completion counters
        }
        return;
    }

    // ===== consumer warps: one row per thread and pass, inputs from shared memory only =====
    constexpr int NCT = NCW * 32;
    double *dst = P.levels[level];
    double *dst2 = NV == 2 ? P.levels2[level] : nullptr;
    const int row_end = P.level_rows[level];
    const bool stream_out = (P.flags & 1) && level == P.k - 1;  // nobody in this launch re-reads the last level
    double dot_acc = 0.0;
    for (int it = 0; it < n_my; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        if (timing && tid == 0) ts[s * 4 + 2] = pk_now();
        const unsigned char *blob = smem + (size_t)s * STAGE_BYTES;
        const int *hdr = reinterpret_cast<const int *>(blob);
        const int row0 = hdr[PKH_ROW0], nrows = hdr[PKH_NROWS], width = hdr[PKH_WIDTH], rp = hdr[PKH_RP];
        const unsigned short *lens = reinterpret_cast<const unsigned short *>(blob + hdr[PKH_OFF_LENS]);
        const unsigned short *lcol = reinterpret_cast<const unsigned short *>(blob + hdr[PKH_OFF_LCOL]);
        const double *val = reinterpret_cast<const double *>(blob + hdr[PKH_OFF_VAL]);
        const double *xb = reinterpret_cast<const double *>(blob + BLOB_CAP);
        for (int rb = 0; rb < nrows; rb += NCT * RPT) {
            int len[RPT];
            double acc[NV][RPT];
#pragma unroll
            for (int q = 0; q < RPT; q++) {
                const int r = rb + q * NCT + tid;
                len[q] = (r < nrows && row0 + r < row_end) ? (int)lens[r] : -1;  // -1: no row
#pragma unroll
                for (int v = 0; v < NV; v++) acc[v][q] = 0.0;
            }
            for (int e0 = 0; e0 < width; e0 += 8) {
                double xv[NV][RPT][8];
#pragma unroll
                for (int q = 0; q < RPT; q++) {
                    const int r = rb + q * NCT + tid;
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (e0 + u < len[q]) {
                            const int cidx = lcol[(e0 + u) * rp + r];
#pragma unroll
                            for (int v = 0; v < NV; v++) xv[v][q][u] = xb[v * XCAP + cidx];
                        }
                }
#pragma unroll
                for (int q = 0; q < RPT; q++) {
                    const int r = rb + q * NCT + tid;
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (e0 + u < len[q]) {
                            const double a = val[(e0 + u) * rp + r];
#pragma unroll
                            for (int v = 0; v < NV; v++) acc[v][q] = row_op<MULADD>(a, xv[v][q][u], acc[v][q]);
                        }
                }
            }
#pragma unroll
            for (int q = 0; q < RPT; q++) {
                const int r = rb + q * NCT + tid;
                if (len[q] >= 0) {
                    if (stream_out) __stcs(dst + row0 + r, acc[0][q]);
                    else dst[row0 + r] = acc[0][q];
                    if (NV == 2) {
                        if (stream_out) __stcs(dst2 + row0 + r, acc[NV - 1][q]);
                        else dst2[row0 + r] = acc[NV - 1][q];
                    }
                    if (NV == 1 && P.dot_w) dot_acc = __fma_rn(P.dot_w[row0 + r], acc[0][q], dot_acc);
                }
            }
        }
        __syncwarp();
        if (timing && tid == 0) ts[s * 4 + 3] = pk_now();
        if (lane == 0) {
            mbar_arrive(&done[s]);                                 // release.cta: the stage may be refilled
            if (P.k > 1) red_release_cta_shared_add(&fin[s], 1u);  // ... and published (the gpu-scope fence is the publisher's)
        }
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
                double sum = 0.0;
                for (int w = 0; w < NCW; w++) sum += red[w];
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
// geometry table
// -----------------------------------------------------------------------------------------------
struct PkVariant {
    int t_rows, blob_cap, xcap, stages, ncw, minb, rpt;
};
//            id T_ROWS BLOB_CAP XCAP STAGES NCW MINB RPT
#define NSK_PK_VARIANTS(X)              \
    X(0, 256, 21504, 1536, 3, 8, 2, 1)  \
    X(1, 256, 21504, 1536, 2, 8, 3, 1)  \
    X(2, 512, 42496, 2560, 3, 8, 1, 2)  \
    X(3, 512, 42496, 2560, 3, 16, 1, 1) \
    X(4, 256, 21504, 1536, 6, 8, 1, 1)  \
    X(5, 128, 11264, 1024, 3, 4, 4, 1)  \
    X(6, 256, 21504, 1536, 2, 4, 3, 2)  \
    X(7, 256, 21504, 1536, 2, 4, 3, 1)  \
    X(8, 256, 21504, 1536, 3, 4, 2, 2)  \
    X(9, 256, 21504, 1536, 6, 4, 1, 2)  \
    X(10, 128, 86016, 2048, 2, 4, 1, 1) \
    X(11, 64, 43008, 1536, 4, 2, 1, 1)  \
    X(12, 64, 43008, 1536, 2, 2, 2, 1)  \
    X(13, 128, 86016, 2048, 2, 8, 1, 1)

static const PkVariant g_pkv[] = {
#define X(id, r, b, x, s, w, m, u) {r, b, x, s, w, m, u},
    NSK_PK_VARIANTS(X)
#undef X
};
static const int g_npkv = sizeof(g_pkv) / sizeof(g_pkv[0]);

typedef void (*pk_fn)(const PkParams);
// Two right-hand sides per launch are built for one geometry: the default short-row one with two stages x two CTAs per
// SM (a stage carries both vectors' x runs).  Other geometries run the vectors one after the other.
constexpr int PK_NV2_VARIANT = 7;
static pk_fn pk_lookup(int variant, bool muladd, int nv, int *smem)
{
    if (nv == 2) {
        if (variant != PK_NV2_VARIANT) return nullptr;
        *smem = (21504 + 2 * 1536 * 8) * 2 + 2 * 2 * 8 + 64 * 8 + 2 * 32 + 2 * 4 + 128;
        return muladd ? packed_kernel<256, 21504, 1536, 2, 4, 2, 1, 2, true> : packed_kernel<256, 21504, 1536, 2, 4, 2, 1, 2, false>;
    }
    switch (variant) {
#define X(id, r, b, x, s, w, m, u)                                          \
    case id:                                                                \
        *smem = (b + x * 8) * s + 2 * s * 8 + 64 * 8 + s * 32 + s * 4 + 128; \
        return muladd ? packed_kernel<r, b, x, s, w, m, u, 1, true> : packed_kernel<r, b, x, s, w, m, u, 1, false>;
        NSK_PK_VARIANTS(X)
#undef X
    }
    return nullptr;
}

static int pk_variant(nsk_csr_t A)
{
    // option value 0 = default; n >= 1 selects table entry n - 1.  Default for short rows: 256-row tiles, 2 stages,
    // 4 consumer warps (one row per thread and pass), 3 CTAs per SM (profiles/r01_sweep_packed_c3.txt).  Rows longer
    // than ~16 nonzeros (FEM operators, 4 dof per node: 58 per row) would fit only ~32 rows in such a stage; they get
    // stages of 84 KB (128 rows x 64 slots) so that a tile keeps enough rows per bulk copy and per consumer pass.
    nsk_ctx_t ctx = A->ctx;
    int v = (int)ctx->opt.packed_variant - 1;
    if (v < 0 || v >= g_npkv) v = A->mean_row > 16.0 ? 10 : 7;
    return v;
}

// -----------------------------------------------------------------------------------------------
// host side: packing (cached per operator and tile geometry), level plans, launches
// -----------------------------------------------------------------------------------------------
struct PkLevelPlan {
    int k = 0, team = 0, lead_pct = 0, bp_global = 0, w0_pct = 0, interleave = 0, l2_pct = 0, nv = 1;
    bool rejected = false;
    std::vector<int> teams;     // CTAs per level, sum <= team * k
    int grid = 0;
    int2 *d_roles = nullptr;
    std::vector<int> level_rows;
    int ngroups = 0, reach = 0, lead = 0;
    std::vector<int> count;
    std::vector<size_t> item_off;
    PkItem *d_items = nullptr;
    int *d_counters = nullptr;
    int *d_group_size = nullptr;
};

struct PackedOp {
    int t_rows = 0, blob_cap = 0, xcap = 0;
    bool ok = false;
    std::string why;
    int ntiles = 0;
    size_t blob_bytes = 0;
    unsigned char *d_blobs = nullptr;
    PkTile *d_tiles = nullptr;
    std::vector<PkTile> h_tiles;
    nsk_tiling csr_view;  // the same tiles as {row0, nrows, nz0, nz1}: input of nsk_wave_deps
    std::vector<PkLevelPlan> plans;
};

// Per-operator state lives in a process-wide map: the map itself is guarded (operators of different contexts may be
// created / destroyed from different host threads); an operator's own entry is used by one thread at a time, like
// the operator (std::map nodes are stable under insertion of other keys).
static std::map<nsk_csr_t, std::vector<PackedOp *>> g_packed;
static std::mutex g_packed_mu;

void nsk_packed_free(nsk_csr_t A)
{
    std::vector<PackedOp *> mine;
    {
        std::lock_guard<std::mutex> lk(g_packed_mu);
        auto it = g_packed.find(A);
        if (it == g_packed.end()) return;
        mine.swap(it->second);
        g_packed.erase(it);
    }
    for (PackedOp *op : mine) {
        for (PkLevelPlan &p : op->plans) {
            if (p.d_items) cudaFree(p.d_items);
            if (p.d_counters) cudaFree(p.d_counters);
            if (p.d_group_size) cudaFree(p.d_group_size);
            if (p.d_roles) cudaFree(p.d_roles);
        }
        if (op->d_blobs) cudaFree(op->d_blobs);
        if (op->d_tiles) cudaFree(op->d_tiles);
        delete op;
    }
}

// Column runs of one tile.  cols: the tile's column indices (any order, duplicates allowed; scratch, sorted in
// place).  A gap of up to GAP unused columns is bridged (a bulk copy per run costs more than 64 idle bytes).
static bool pk_segments(std::vector<int> &cols, int n_cols, int xcap, PkTile &t)
{
    constexpr int GAP = 8;
    t.nseg = 0;
    t.xlen = 0;
    t.tail = 0;
    for (int s = 0; s < PK_MAXSEG; s++) { t.seg_start[s] = 0; t.seg_lenoff[s] = 0; }
    if (cols.empty()) return true;
    std::sort(cols.begin(), cols.end());
    int s0 = cols[0] & ~1, s1 = cols[0] + 1;  // current run [s0, s1)
    auto flush = [&]() {
        int e = (s1 + 1) & ~1;
        bool tail = false;
        if (e > n_cols) {  // only possible when n_cols is odd and the run reaches column n_cols - 1
            e = n_cols - 1;   // even: the run stops one short, the last column gets its own slot after it
            tail = true;
        }
        const int len = e - s0;
        if (len > 0) {
            if (t.nseg >= PK_MAXSEG || t.xlen + len > xcap || len > 0xffff || t.xlen > 0xffff) return false;
            t.seg_start[t.nseg] = s0;
            t.seg_lenoff[t.nseg] = len | (t.xlen << 16);
            t.nseg++;
            t.xlen += len;
        }
        if (tail) {
            if (t.xlen + 2 > xcap) return false;
            t.tail = t.xlen + 1;  // slot t.xlen, stored + 1 (0 = none); the bulk copies still move t.xlen doubles
        }
        return true;
    };
    for (size_t i = 1; i < cols.size(); i++) {
        const int cidx = cols[i];
        if (cidx < s1) continue;
        if (cidx <= ((s1 + 1) & ~1) + GAP) { s1 = cidx + 1; continue; }
        if (!flush()) return false;
        s0 = cidx & ~1;
        s1 = cidx + 1;
    }
    return flush();
}

// Host-only packer: everything about the packed format that does not need a GPU (also reachable through
// nsk_pack_host_* for the CPU test-suite).  Returns "" on success, else why the operator does not pack.
struct PackedHost {
    std::vector<nsk_tile> tiles;
    std::vector<PkTile> ptiles;
    std::vector<unsigned char> blobs;  // + 64 bytes of slack
    size_t blob_bytes = 0;
};

static std::string pk_pack_host(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                                const std::vector<int> &breaks, int t_rows, int blob_cap, int xcap, PackedHost &out)
{
    if (n == 0 || nnz == 0) return "empty operator";
    // 1. tiles: up to t_rows consecutive rows, never across a break, blob within the stage
    std::vector<nsk_tile> &tiles = out.tiles;
    std::vector<int> widths;
    tiles.clear();
    {
        size_t bi = 0;
        int r = 0;
        while (r < n) {
            while (bi < breaks.size() && breaks[bi] <= r) bi++;
            const int seg_end = bi < breaks.size() ? std::min(n, breaks[bi]) : n;
            int rows = std::min(t_rows, seg_end - r);
            int width = 0;
            for (;;) {
                width = 0;
                for (int i = r; i < r + rows; i++) width = std::max(width, ptrow[i + 1] - ptrow[i]);
                if (pk_blob_bytes(rows, width) <= blob_cap) break;
                if (rows == 1) return "a row is longer than a stage";
                rows = rows / 2;
            }
            tiles.push_back(nsk_tile{r, rows, ptrow[r], ptrow[r + rows]});
            widths.push_back(width);
            r += rows;
        }
    }
    const int ntiles = (int)tiles.size();
    {
        size_t explicit_bytes = 0;
        for (int t = 0; t < ntiles; t++) explicit_bytes += (size_t)pk_blob_bytes(tiles[t].nrows, widths[t]);
        const double csr_equiv = 10.0 * (double)nnz + 2.0 * n;
        if ((double)explicit_bytes > 1.35 * csr_equiv + 65536.0) return "row lengths too ragged for slot-major tiles";
    }

    std::vector<PkTile> &ptiles = out.ptiles;
    ptiles.assign(ntiles, PkTile());
    std::atomic<int> failed(0);
    struct Runs {
        int nseg, start[PK_MAXSEG], end[PK_MAXSEG], off[PK_MAXSEG];
        explicit Runs(const PkTile &pt) : nseg(pt.nseg)
        {
            for (int sg = 0; sg < nseg; sg++) {
                start[sg] = pt.seg_start[sg];
                end[sg] = start[sg] + (pt.seg_lenoff[sg] & 0xffff);
                off[sg] = (pt.seg_lenoff[sg] >> 16) & 0xffff;
            }
        }
        // local column of global column cidx; sg = the run the previous column fell into (columns mostly ascend)
        int local(int cidx, int &sg) const
        {
            if (sg >= nseg || cidx < start[sg] || cidx >= end[sg]) {
                sg = 0;
                while (sg < nseg && !(cidx >= start[sg] && cidx < end[sg])) sg++;
            }
            return off[sg] + (cidx - start[sg]);
        }
    };
    auto local_col = [&](const PkTile &pt, const Runs &R, int cidx, int &sg) {
        if (pt.tail && cidx == n_cols - 1) return pt.tail - 1;  // the hand-copied last element of an odd-length vector
        return R.local(cidx, sg);
    };
    auto parallel = [&](const std::function<void(int)> &per_tile) {
        std::atomic<int> next(0);
        auto work = [&]() {
            for (;;) {
                const int t0 = next.fetch_add(64);
                if (t0 >= ntiles || failed.load()) return;
                for (int t = t0; t < std::min(ntiles, t0 + 64); t++) per_tile(t);
            }
        };
        const int nth = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (int i = 1; i < nth; i++) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
    };

    // 2. per tile: column runs
    parallel([&](int t) {
        const nsk_tile &tl = tiles[t];
        PkTile &pt = ptiles[t];
        std::vector<int> cols(indcol + tl.nz0, indcol + tl.nz1);
        if (!pk_segments(cols, n_cols, xcap, pt)) { failed.store(1); return; }
        pt.row0 = tl.row0; pt.nrows = tl.nrows;
    });
    if (failed.load()) return "a tile references x in too many / too long runs";

    std::vector<size_t> off(ntiles + 1, 0);
    for (int t = 0; t < ntiles; t++)
        off[t + 1] = off[t] + (size_t)pk_blob_bytes(tiles[t].nrows, widths[t]);

    // 3. the blobs, slot-major
    std::vector<unsigned char> &blobs = out.blobs;
    blobs.assign(off[ntiles] + 64, 0);
    parallel([&](int t) {
        const nsk_tile &tl = tiles[t];
        PkTile &pt = ptiles[t];
        const int width = widths[t], rp = pk_round_up(tl.nrows, 32);
        pt.blob_off = (long long)off[t];
        pt.blob_bytes = (int)(off[t + 1] - off[t]);
        unsigned char *b = blobs.data() + off[t];
        int *hdr = reinterpret_cast<int *>(b);
        const int off_base = PKH_WORDS * 4;
        const int off_lens = off_base;
        const int off_lcol = off_lens + 2 * rp;
        const int off_val = off_lcol + 2 * width * rp;
        hdr[PKH_ROW0] = tl.row0; hdr[PKH_NROWS] = tl.nrows; hdr[PKH_WIDTH] = width; hdr[PKH_RP] = rp;
        hdr[PKH_OFF_LENS] = off_lens; hdr[PKH_OFF_LCOL] = off_lcol; hdr[PKH_OFF_VAL] = off_val;
        hdr[PKH_FORMAT] = 0; hdr[PKH_OFF_BASE] = 0;
        unsigned short *lens = reinterpret_cast<unsigned short *>(b + off_lens);
        unsigned short *lcol = reinterpret_cast<unsigned short *>(b + off_lcol);
        double *val = reinterpret_cast<double *>(b + off_val);
        const Runs R(pt);
        for (int r = 0; r < tl.nrows; r++) {
            const int p = ptrow[tl.row0 + r], q = ptrow[tl.row0 + r + 1];
            int sg = 0;
            lens[r] = (unsigned short)(q - p);
            for (int j = p; j < q; j++) {
                lcol[(size_t)(j - p) * rp + r] = (unsigned short)local_col(pt, R, indcol[j], sg);
                val[(size_t)(j - p) * rp + r] = coef[j];
            }
        }
    });
    if (failed.load()) return "a tile references x in too many / too long runs";
    out.blob_bytes = off[ntiles];
    return "";
}

static PackedOp *pk_get(nsk_csr_t A, const PkVariant &V)
{
    g_packed_mu.lock();
    std::vector<PackedOp *> &ops = g_packed[A];
    g_packed_mu.unlock();
    for (PackedOp *op : ops)
        if (op->t_rows == V.t_rows && op->blob_cap == V.blob_cap && op->xcap == V.xcap) return op;
    PackedOp *op = new PackedOp();
    op->t_rows = V.t_rows; op->blob_cap = V.blob_cap; op->xcap = V.xcap;
    ops.push_back(op);
    const int n = A->n;
    const std::vector<int> &ptrow = nsk_csr_host_ptrow(A);
    if (n == 0 || A->nnz == 0) { op->why = "empty operator"; return op; }
    if ((int)ptrow.size() != n + 1) { op->why = "host row pointers missing"; return op; }
    // cheap refusals first (row lengths only), then the operator's entries: the caller's host arrays are gone, read
    // them back once
    std::vector<int> indcol((size_t)A->nnz);
    std::vector<double> coef((size_t)A->nnz);
    if (cudaMemcpy(indcol.data(), A->d_indcol, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(coef.data(), A->d_coef, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost) != cudaSuccess) {
        op->why = "reading the operator back failed";
        return op;
    }
    PackedHost H;
    op->why = pk_pack_host(n, A->n_cols, A->nnz, ptrow.data(), indcol.data(), coef.data(), A->breaks, V.t_rows, V.blob_cap,
                           V.xcap, H);
    if (!op->why.empty()) return op;
    const int ntiles = (int)H.tiles.size();
    if (cudaMalloc(&op->d_blobs, H.blobs.size()) != cudaSuccess ||
        cudaMalloc(&op->d_tiles, sizeof(PkTile) * (size_t)ntiles + 128) != cudaSuccess) {
        op->why = "allocation of the packed operator failed";
        cudaGetLastError();
        return op;
    }
    cudaMemcpy(op->d_blobs, H.blobs.data(), H.blobs.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(op->d_tiles, H.ptiles.data(), sizeof(PkTile) * (size_t)ntiles, cudaMemcpyHostToDevice);
    op->ntiles = ntiles;
    op->blob_bytes = H.blob_bytes;
    op->h_tiles.swap(H.ptiles);
    op->csr_view.tile_rows = V.t_rows;
    op->csr_view.ntiles = ntiles;
    op->csr_view.nlong = 0;
    op->csr_view.h_tiles.swap(H.tiles);
    op->ok = true;
    return op;
}

// Host-only: the per-level item lists, the group sizes the completion counters are compared with, and the CTA role
// table.  teams[] comes in as the requested team sizes and goes out clipped to the items a level really has.
struct PkSchedule {
    std::vector<PkItem> items;       // all levels back to back
    std::vector<size_t> item_off;    // per level
    std::vector<int> count;          // per level
    std::vector<int> gsize;          // [k][ngroups]
    std::vector<int2> roles;         // per CTA {level, index in team}
};

static void pk_build_schedule(const std::vector<PkTile> &tiles, const std::vector<int> &pos_tile, const std::vector<int> &ghi,
                              const std::vector<int> &lr, int k, int ngroups, int lead, int bp_global, int interleave,
                              std::vector<int> &teams, PkSchedule &S)
{
    const int ntiles = (int)tiles.size();
    S.items.clear();
    S.items.reserve((size_t)k * ntiles);
    S.gsize.assign((size_t)k * ngroups, 0);
    S.count.assign(k, 0);
    S.item_off.assign(k, 0);
    for (int l = 0; l < k; l++) {
        S.item_off[l] = S.items.size();
        for (int pos = 0; pos < ntiles; pos++) {
            const int t = pos_tile[pos];
            const PkTile &pt = tiles[t];
            if (pt.row0 >= lr[l]) continue;  // outside this level's row prefix (distributed shrink)
            int gback = -1;
            // adjacent mode: level l trails l+1 by at most `lead`; global mode: level 0 trails k-1 by (k-1)*lead
            const int hold = bp_global ? (l == 0 && k > 1 ? (k - 1) * lead : -1) : (l < k - 1 ? lead : -1);
            if (hold >= 0 && pos - hold >= 0) gback = (pos - hold) / WF_GROUP - 1;
            S.items.push_back(PkItem{t, pos, ghi[t], gback, pt.blob_off, pt.blob_bytes, 0});
            S.gsize[(size_t)l * ngroups + pos / WF_GROUP]++;
            S.count[l]++;
        }
    }
    // role table: levels interleaved in proportion to their team sizes (consecutive block indices land on different
    // SMs, so every SM hosts a mix of levels), or level by level
    for (int l = 0; l < k; l++) teams[l] = std::max(1, std::min(teams[l], std::max(1, S.count[l])));
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

// CPU model of the kernel's protocol, for the test-suite: every CTA runs its items strictly in order (the dependency
// warp is head-of-line blocking), an item may start when all groups <= ghi of level l-1 and all groups <= gback of the
// level that holds it back have reached their group sizes, and it reports to its own group when it finishes; CTAs are
// visited in a seeded random order and up to `stages` items per CTA may be open at once (they finish in random order).
// Returns the number of items that completed; the schedule is sound iff that equals the total.
// reads: per tile, the tiles whose rows its nonzeros really reference (exact, from the local columns); ntile_done marks
// (level, tile) complete.  *violations counts items opened while a tile they read was not complete at the level below
// (and inside that level's row prefix): the declared dependencies (prefix of groups <= ghi) must imply data readiness.
static long long pk_simulate(const PkSchedule &S, const std::vector<int> &teams, int k, int ngroups, int bp_global, int stages,
                             unsigned seed, const std::vector<std::vector<int>> *reads = nullptr,
                             const std::vector<int> *tile_row0 = nullptr, const std::vector<int> *lr = nullptr,
                             long long *violations = nullptr)
{
    const int grid = (int)S.roles.size();
    const int ntiles_all = reads ? (int)reads->size() : 0;
    std::vector<std::vector<char>> tile_done(reads ? k : 0, std::vector<char>((size_t)ntiles_all, 0));
    long long bad = 0;
    std::vector<std::vector<int>> cnt(k, std::vector<int>(ngroups, 0));
    std::vector<int> water(k, 0);  // all groups < water[l] complete at level l
    auto advance = [&](int l) {
        while (water[l] < ngroups && cnt[l][water[l]] >= S.gsize[(size_t)l * ngroups + water[l]]) water[l]++;
    };
    for (int l = 0; l < k; l++) advance(l);
    std::vector<int> next(grid, 0);                 // next item index (within the CTA's own sequence) to open
    std::vector<std::vector<int>> open(grid);       // opened, not yet finished (global item indices)
    long long done = 0, total = (long long)S.items.size();
    unsigned rng = seed * 2654435761u + 12345u;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    std::vector<int> order(grid);
    for (int b = 0; b < grid; b++) order[b] = b;
    bool progress = true, force = false;
    while (done < total && (progress || !force)) {
        force = !progress;  // a pass in which every coin said "not yet" is not a deadlock: the next pass finishes items
        progress = false;
        for (int i = grid - 1; i > 0; i--) std::swap(order[i], order[rnd() % (unsigned)(i + 1)]);
        for (int oi = 0; oi < grid; oi++) {
            const int b = order[oi];
            const int level = S.roles[b].x, c = S.roles[b].y, G = teams[level];
            // finish one open item (random pick) with probability 1/2, or whenever the ring is full
            if (!open[b].empty() && (force || (rnd() & 1) || (int)open[b].size() >= stages)) {
                const int pick = (int)(rnd() % (unsigned)open[b].size());
                const PkItem &it = S.items[(size_t)open[b][pick]];
                open[b].erase(open[b].begin() + pick);
                cnt[level][it.pos / WF_GROUP]++;
                if (reads) tile_done[level][(size_t)it.tile] = 1;
                advance(level);
                done++;
                progress = true;
            }
            // open the next item if the ring has room and its inputs are complete
            const long long idx = (long long)c + (long long)next[b] * G;
            if ((int)open[b].size() < stages && idx < S.count[level]) {
                const PkItem &it = S.items[S.item_off[level] + (size_t)idx];
                const int lb = bp_global ? k - 1 : level + 1;
                const bool fwd_ok = level == 0 || water[level - 1] > it.ghi || water[level - 1] >= ngroups;
                const bool back_ok = it.gback < 0 || water[lb] > it.gback || water[lb] >= ngroups;
                if (fwd_ok && back_ok) {
                    if (reads && level > 0)
                        for (int d : (*reads)[(size_t)it.tile])
                            if ((*tile_row0)[(size_t)d] < (*lr)[level - 1] && !tile_done[level - 1][(size_t)d]) bad++;
                    open[b].push_back((int)(S.item_off[level] + (size_t)idx));
                    next[b]++;
                    progress = true;
                }
            }
        }
    }
    if (violations) *violations = bad;
    return done;
}

// ---- host-only access to the packer (CPU tests: pack, then expand the blobs back to CSR and compare) --------------
struct nsk_packed_host_s {
    PackedHost H;
    std::string why;
    int n = 0, n_cols = 0;
};

static int pk_host_create(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                          int variant, void **out)
{
    if (!out || !ptrow || (nnz > 0 && (!indcol || !coef))) return NSK_ERR_INVALID;
    if (variant < 0 || variant >= g_npkv) return NSK_ERR_INVALID;
    nsk_packed_host_s *h = new nsk_packed_host_s();
    h->n = n;
    h->n_cols = n_cols;
    const PkVariant &V = g_pkv[variant];
    h->why = pk_pack_host(n, n_cols, nnz, ptrow, indcol, coef, std::vector<int>(), V.t_rows, V.blob_cap, V.xcap, h->H);
    *out = h;
    return NSK_OK;
}

NSK_API int nsk_pack_host_create(int n, int n_cols, int64_t nnz, const int *ptrow, const int *indcol, const double *coef,
                                 int variant, void **out)
{
    return pk_host_create(n, n_cols, nnz, ptrow, indcol, coef, variant, out);
}

NSK_API const char *nsk_pack_host_why(void *handle) { return static_cast<nsk_packed_host_s *>(handle)->why.c_str(); }

NSK_API int64_t nsk_pack_host_bytes(void *handle)
{
    nsk_packed_host_s *h = static_cast<nsk_packed_host_s *>(handle);
    return h->why.empty() ? (int64_t)h->H.blob_bytes : 0;
}

// Expands the packed operator back to CSR (row pointers, GLOBAL column indices through the tiles' runs, values) and
// reports the largest number of runs / x doubles any tile needs.  Arrays sized n+1 / nnz by the caller.
NSK_API int nsk_pack_host_expand(void *handle, int *ptrow, int *indcol, double *coef, int *max_runs, int *max_xlen)
{
    nsk_packed_host_s *h = static_cast<nsk_packed_host_s *>(handle);
    if (!h->why.empty()) return NSK_ERR_UNSUPPORTED;
    int mr = 0, mx = 0;
    int64_t k = 0;
    ptrow[0] = 0;
    for (size_t t = 0; t < h->H.tiles.size(); t++) {
        const PkTile &pt = h->H.ptiles[t];
        mr = std::max(mr, pt.nseg);
        mx = std::max(mx, pt.xlen);
        const PkBlobView B(h->H.blobs.data() + pt.blob_off);
        const int *hdr = B.hdr;
        for (int r = 0; r < hdr[PKH_NROWS]; r++) {
            for (int e = 0; e < B.len(r); e++) {
                const int lc = B.col(e, r);
                int g = -1;
                if (pt.tail && lc == pt.tail - 1) g = h->n_cols - 1;
                for (int s = 0; s < pt.nseg && g < 0; s++) {
                    const int len = pt.seg_lenoff[s] & 0xffff, xoff = (pt.seg_lenoff[s] >> 16) & 0xffff;
                    if (lc >= xoff && lc < xoff + len) { g = pt.seg_start[s] + (lc - xoff); break; }
                }
                if (g < 0) return NSK_ERR_INVALID;
                indcol[k] = g;
                coef[k] = B.value(e, r);
                k++;
            }
            ptrow[hdr[PKH_ROW0] + r + 1] = (int)k;
        }
    }
    if (max_runs) *max_runs = mr;
    if (max_xlen) *max_xlen = mx;
    return NSK_OK;
}

// Builds the level schedule for a packed operator exactly like the GPU path (dependencies from the tiles' column
// extents, natural row order, optional per-level row prefixes) and runs the CPU protocol model.  Returns the number
// of items that did NOT complete (0 = sound), or a negative status.  *items_out receives the total item count.
NSK_API long long nsk_pack_host_simulate(void *handle, int k, int lead_slack_tiles, int resident, int w0_pct, int bp_global,
                                         int interleave, int stages, const int *level_rows, unsigned seed,
                                         long long *items_out, int *reach_out, int ghi_bias)
{
    nsk_packed_host_s *h = static_cast<nsk_packed_host_s *>(handle);
    if (!h->why.empty()) return NSK_ERR_UNSUPPORTED;
    if (k < 1 || k > NSK_MAX_K || resident < k || stages < 1) return NSK_ERR_INVALID;
    const int ntiles = (int)h->H.tiles.size();
    const int ngroups = (ntiles + WF_GROUP - 1) / WF_GROUP;
    // dependencies: the tile range covered by a tile's x runs (columns >= n are ghost entries of x, level 0 only)
    std::vector<int> pos_tile(ntiles), ghi(ntiles, 0), row0s(ntiles);
    for (int t = 0; t < ntiles; t++) { pos_tile[t] = t; row0s[t] = h->H.tiles[t].row0; }
    int reach = 0;
    for (int t = 0; t < ntiles; t++) {
        const PkTile &pt = h->H.ptiles[t];
        int mx = -1;
        for (int s = 0; s < pt.nseg; s++) {
            int last = pt.seg_start[s] + (pt.seg_lenoff[s] & 0xffff) - 1;
            if (last >= h->n) last = h->n - 1;
            mx = std::max(mx, last);
        }
        if (pt.tail) mx = std::max(mx, std::min(h->n_cols, h->n) - 1);
        if (mx < 0) mx = pt.row0;
        const int tmax = (int)(std::upper_bound(row0s.begin(), row0s.end(), mx) - row0s.begin()) - 1;
        ghi[t] = std::max(0, tmax) / WF_GROUP;
        reach = std::max(reach, std::min(ntiles - 1, (ghi[t] + 1) * WF_GROUP - 1) - t);
        ghi[t] = std::max(0, ghi[t] + ghi_bias);  // tests weaken the dependencies on purpose (ghi_bias < 0) to see the model object
    }
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : h->n;
    std::vector<int> teams(k, 0);
    {
        const double wsum = (double)w0_pct + 100.0 * (k - 1);
        int used = 0;
        for (int l = 1; l < k; l++) { teams[l] = std::max(1, (int)(resident * 100.0 / wsum)); used += teams[l]; }
        teams[0] = std::max(1, resident - used);
    }
    // a negative slack undercuts the safe minimum on purpose: the test-suite checks that the model then reports a deadlock
    const int lead = std::max(1, reach + 1 + WF_GROUP + lead_slack_tiles);
    PkSchedule S;
    pk_build_schedule(h->H.ptiles, pos_tile, ghi, lr, k, ngroups, lead, bp_global, interleave, teams, S);
    // exact read sets: every local column of every row, mapped back through the runs to a global column, then to a tile
    std::vector<std::vector<int>> reads((size_t)ntiles);
    for (int t = 0; t < ntiles; t++) {
        const PkTile &pt = h->H.ptiles[t];
        const PkBlobView B(h->H.blobs.data() + pt.blob_off);
        const int *hdr = B.hdr;
        std::vector<int> &rd = reads[(size_t)t];
        for (int r = 0; r < hdr[PKH_NROWS]; r++)
            for (int e = 0; e < B.len(r); e++) {
                const int lc = B.col(e, r);
                int g = -1;
                if (pt.tail && lc == pt.tail - 1) g = h->n_cols - 1;
                for (int sgm = 0; sgm < pt.nseg && g < 0; sgm++) {
                    const int len = pt.seg_lenoff[sgm] & 0xffff, xoff = (pt.seg_lenoff[sgm] >> 16) & 0xffff;
                    if (lc >= xoff && lc < xoff + len) g = pt.seg_start[sgm] + (lc - xoff);
                }
                if (g < 0 || g >= h->n) continue;  // ghost entry of x: read by level 0 only
                const int d = (int)(std::upper_bound(row0s.begin(), row0s.end(), g) - row0s.begin()) - 1;
                if (rd.empty() || rd.back() != d) rd.push_back(d);
            }
        std::sort(rd.begin(), rd.end());
        rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
    }
    long long violations = 0;
    const long long done = pk_simulate(S, teams, k, ngroups, bp_global, stages, seed, &reads, &row0s, &lr, &violations);
    if (items_out) *items_out = (long long)S.items.size();
    if (reach_out) *reach_out = reach;
    if (violations > 0) return -1000000 - violations;  // a dependency hole: an item would read rows not yet produced
    return (long long)S.items.size() - done;
}

NSK_API void nsk_pack_host_destroy(void *handle) { delete static_cast<nsk_packed_host_s *>(handle); }

static int pk_launch_shape(nsk_ctx_t ctx, int variant, bool muladd, int k, int nv, pk_fn *fn_out, int *smem_out, int *team)
{
    const PkVariant &V = g_pkv[variant];
    int smem = 0;
    pk_fn fn = pk_lookup(variant, muladd, nv, &smem);
    if (!fn) {
        nsk_set_error(ctx, "packed path: no two-vector kernel for this tile geometry");
        return NSK_ERR_UNSUPPORTED;
    }
    NSK_REQUIRE(ctx, smem <= (int)ctx->prop.sharedMemPerBlockOptin, "packed kernel stage ring exceeds shared memory");
    NSK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    NSK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ((nv == 2 ? 4 : V.ncw) + 3) * 32, smem));
    if (ctx->opt.spmv_ctas_per_sm > 0) per_sm = std::min(per_sm, (int)ctx->opt.spmv_ctas_per_sm);
    const int resident = ctx->prop.multiProcessorCount * per_sm;
    *team = resident;  // all resident CTAs; the plan shares them out over the levels
    *fn_out = fn;
    *smem_out = smem;
    return NSK_OK;
}

static PkLevelPlan *pk_level_plan(nsk_csr_t A, PackedOp *op, int k, const int *level_rows, int resident, int nv, const char **why)
{
    nsk_ctx_t ctx = A->ctx;
    const int team = resident / k;  // even share: the unit of the slack below
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : A->n;
    const int lead_pct = (int)ctx->opt.wave_slack_pct;  // < 0: size the window from the L2 budget
    const int bp_global = ctx->opt.pipe_bp_global >= 0 ? (ctx->opt.pipe_bp_global ? 1 : 0) : 1;
    const int w0_pct = k > 1 ? (ctx->opt.pipe_w0_pct > 0 ? (int)ctx->opt.pipe_w0_pct : 100) : 100;
    const int interleave = ctx->opt.pipe_interleave ? 1 : 0;
    const int l2_pct = (int)ctx->opt.wave_l2_pct;  // part of the key: it sizes the window and decides refusals
    for (PkLevelPlan &p : op->plans)
        if (p.k == k && p.team == resident && p.level_rows == lr && p.lead_pct == lead_pct && p.bp_global == bp_global &&
            p.w0_pct == w0_pct && p.interleave == interleave && p.l2_pct == l2_pct && p.nv == nv) {
            if (p.rejected) { *why = "wavefront window exceeds the L2 budget"; return nullptr; }
            return &p;
        }
    const int ntiles = op->ntiles;
    PkLevelPlan p;
    p.k = k; p.team = resident; p.lead_pct = lead_pct; p.level_rows = lr; p.bp_global = bp_global;
    p.w0_pct = w0_pct; p.interleave = interleave; p.l2_pct = l2_pct; p.nv = nv;
    // teams: `team * k` resident CTAs shared out with weight w0 for level 0 (it streams from HBM: longer fills, so
    // it needs more stages in flight for the same rate) and 100 for every other level
    {
        const double wsum = (double)w0_pct + 100.0 * (k - 1);
        p.teams.assign(k, 0);
        int used = 0;
        for (int l = 1; l < k; l++) {
            p.teams[l] = std::max(1, (int)(resident * 100.0 / wsum));
            used += p.teams[l];
        }
        p.teams[0] = std::max(1, resident - used);
    }
    std::vector<int> pos_tile(ntiles), ghi(ntiles, 0);
    int ngroups = (ntiles + WF_GROUP - 1) / WF_GROUP;
    if (k > 1) {
        WaveDeps D;
        if (!nsk_wave_deps(A, op->csr_view, D, why)) return nullptr;
        pos_tile = D.tile_at_pos;
        ghi = D.ghi;
        ngroups = D.ngroups;
        p.reach = D.reach;
        // lead = how far a level may run ahead of the level that holds it back, per hop: at least the pattern's
        // reach + one completion group, plus slack that absorbs the publish -> poll -> copy latency (several us,
        // i.e. hundreds of tiles at full rate).  The window that must stay in L2 is (k-1)*lead tiles of matrix
        // data plus the level vectors over it; by default the slack is whatever the L2 budget allows -- measured
        // on 256^3: time falls with the window until it reaches ~100 MB, then HBM re-reads set in.
        const double tile_bytes = (double)op->blob_bytes / ntiles + 8.0 * op->t_rows * (k + 1) * nv;
        // budget: 70 % of L2 by default -- ncu (profiles/r01_ncu_packed_dram_traffic.txt): an 81 MB window costs the
        // compulsory 1.47 GB of HBM reads, a 101 MB one 4.9 GB (the level vectors and x runs share the cache)
        const double budget = (ctx->opt.wave_l2_pct > 0 ? (double)ctx->opt.wave_l2_pct : 70.0) / 100.0 *
                              (double)ctx->prop.l2CacheSize;
        const int lead_min = D.reach + 1 + WF_GROUP;
        if (lead_pct >= 0)
            p.lead = lead_min + (int)((double)lead_pct / 100.0 * 2.0 * team + 0.999);  // unit: two rounds of an even team
        else
            p.lead = std::max(lead_min + 2 * WF_GROUP, (int)(budget / ((double)(k - 1) * tile_bytes)));
        const double window = (double)(k - 1) * p.lead * tile_bytes;
        // A window that binds (lead < ntiles) but leaves little slack beyond the pattern's reach makes every level wait
        // on the loop latency: measured on a 512 x 512 x 64 slab, three levels with 589 tiles of slack per hop cost
        // 0.25 ms per level -- more than separate products (0.215) -- while two levels with 2460 cost 0.19.  Below
        // ~180 k rows of slack the planner refuses and the caller fuses fewer levels per launch instead.
        const int min_slack = 180000 / std::max(1, op->t_rows);
        const bool thin = lead_pct < 0 && p.lead < ntiles && p.lead - lead_min < min_slack;
        if (window > budget * 1.0001 || thin) {
            *why = "wavefront window exceeds the L2 budget";
            p.rejected = true;
            op->plans.push_back(p);
            return nullptr;
        }
    } else {
        for (int t = 0; t < ntiles; t++) pos_tile[t] = t;
    }
    p.ngroups = ngroups;
    PkSchedule S;
    pk_build_schedule(op->h_tiles, pos_tile, ghi, lr, k, ngroups, p.lead, bp_global, interleave, p.teams, S);
    p.count = S.count;
    p.item_off = S.item_off;
    p.grid = (int)S.roles.size();
    std::vector<PkItem> &items = S.items;
    std::vector<int> &gsize = S.gsize;
    std::vector<int2> &roles = S.roles;
    if (cudaMalloc(&p.d_roles, sizeof(int2) * roles.size()) != cudaSuccess ||
        cudaMalloc(&p.d_items, sizeof(PkItem) * (items.size() + 1)) != cudaSuccess ||
        cudaMalloc(&p.d_counters, sizeof(int) * ((size_t)k * ngroups + 4)) != cudaSuccess ||
        cudaMalloc(&p.d_group_size, sizeof(int) * (size_t)k * ngroups) != cudaSuccess) {
        *why = "plan allocation failed";
        cudaGetLastError();
        return nullptr;
    }
    cudaMemcpy(p.d_items, items.data(), sizeof(PkItem) * items.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_group_size, gsize.data(), sizeof(int) * gsize.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_roles, roles.data(), sizeof(int2) * roles.size(), cudaMemcpyHostToDevice);
    op->plans.push_back(p);
    return &op->plans.back();
}

static bool pk_aligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Returns NSK_ERR_UNSUPPORTED (and sets the context's error text) when the packed path does not apply; the
// callers then run the CSR kernels.
static int pk_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2, double *const *d_levels2,
                  nsk_mode mode, const int *level_rows, const double *dot_w, int dot_slot);

int nsk_packed_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode, const int *level_rows,
                   const double *dot_w, int dot_slot)
{
    return pk_run(A, k, d_x, d_levels, nullptr, nullptr, mode, level_rows, dot_w, dot_slot);
}

// Two right-hand sides through ONE launch: levels[l] = A^(l+1) x and levels2[l] = A^(l+1) x2; every tile's blob is
// streamed once for both (s-step Krylov methods need the powers of p and of r in the same outer step).
int nsk_packed_run2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                    double *const *d_levels2, nsk_mode mode, const int *level_rows)
{
    return pk_run(A, k, d_x, d_levels, d_x2, d_levels2, mode, level_rows, nullptr, -1);
}

static int pk_run(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2, double *const *d_levels2,
                  nsk_mode mode, const int *level_rows, const double *dot_w, int dot_slot)
{
    nsk_ctx_t ctx = A->ctx;
    const int nv = d_x2 ? 2 : 1;
    const int variant = nv == 2 && ctx->opt.packed_variant <= 0 && A->mean_row <= 16.0 ? PK_NV2_VARIANT : pk_variant(A);
    const PkVariant &V = g_pkv[variant];
    if (!pk_aligned(d_x) || (d_x2 && !pk_aligned(d_x2))) { nsk_set_error(ctx, "packed path: x is not 16-byte aligned"); return NSK_ERR_UNSUPPORTED; }
    for (int l = 0; l < k; l++)
        if (!pk_aligned(d_levels[l]) || (d_levels2 && !pk_aligned(d_levels2[l]))) {
            nsk_set_error(ctx, "packed path: output is not 16-byte aligned");
            return NSK_ERR_UNSUPPORTED;
        }
    PackedOp *op = pk_get(A, V);
    if (!op->ok) { nsk_set_error(ctx, "packed path not applicable: %s", op->why.c_str()); return NSK_ERR_UNSUPPORTED; }
    pk_fn fn; int smem = 0, team = 0;
    NSK_TRY(pk_launch_shape(ctx, variant, mode == NSK_EXACT_MULADD, k, nv, &fn, &smem, &team));
    if (team < k) { nsk_set_error(ctx, "packed path: fewer resident CTAs than levels"); return NSK_ERR_UNSUPPORTED; }
    const char *why = "";
    PkLevelPlan *plan = pk_level_plan(A, op, k, level_rows, team, nv, &why);
    if (!plan) { nsk_set_error(ctx, "packed matrix powers not applicable: %s", why); return NSK_ERR_UNSUPPORTED; }
    int maxcount = 0;
    for (int l = 0; l < k; l++) maxcount = std::max(maxcount, plan->count[l]);
    if (maxcount == 0) return NSK_OK;
    if (dot_w) NSK_REQUIRE(ctx, k == 1 && plan->grid <= NSK_MAX_PARTIALS, "fused dot: k = 1 and a bounded grid");

    if (k > 1)
        NSK_CUDA(ctx, cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * ((size_t)k * plan->ngroups + 4), ctx->stream));
    PkParams P;
    for (int l = 0; l < NSK_MAX_K; l++) {
        P.items[l] = l < k ? plan->d_items + plan->item_off[l] : nullptr;
        P.count[l] = l < k ? plan->count[l] : 0;
        P.levels[l] = l < k ? d_levels[l] : nullptr;
        P.levels2[l] = l < k && d_levels2 ? d_levels2[l] : nullptr;
        P.level_rows[l] = l < k ? plan->level_rows[l] : 0;
    }
    P.x2 = d_x2;
    P.tiles = op->d_tiles;
    P.blobs = op->d_blobs;
    P.counters = plan->d_counters;
    P.group_size = plan->d_group_size;
    P.ngroups = plan->ngroups;
    P.n_cols = A->n_cols;
    P.x = d_x;
    P.k = k;
    for (int l = 0; l < NSK_MAX_K; l++) P.team[l] = l < k ? plan->teams[l] : 0;
    P.cta_role = plan->d_roles;
    P.bp_global = plan->bp_global;
    P.flags = (int)ctx->opt.pk_flags;
    P.timing = nullptr;
    if (ctx->opt.pk_timing) {
        void *tb = nullptr;
        NSK_TRY(nsk_stage(ctx, 7, sizeof(unsigned long long) * 16 * (size_t)plan->grid, &tb));
        NSK_CUDA(ctx, cudaMemsetAsync(tb, 0, sizeof(unsigned long long) * 16 * (size_t)plan->grid, ctx->stream));
        P.timing = reinterpret_cast<unsigned long long *>(tb);
    }
    P.dot_w = dot_w;
    P.partials = ctx->d_partials;
    P.ticket = ctx->d_ticket;
    P.dot_out = dot_w ? ctx->d_scalars + dot_slot : nullptr;
    if (k > 1) {
        // CTAs of different levels wait on each other: co-residency must be guaranteed, not assumed from the occupancy
        // calculator (another kernel on the device, MPS or a user stream could hold SMs)
        void *args[] = {&P};
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)fn, dim3(plan->grid), dim3(((nv == 2 ? 4 : V.ncw) + 3) * 32),
                                                    args, (size_t)smem, ctx->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
            cudaGetLastError();
            nsk_set_error(ctx, "packed matrix powers: the grid cannot be made co-resident on this device now");
            return NSK_ERR_UNSUPPORTED;
        }
        NSK_CUDA(ctx, e);
    } else {
        fn<<<plan->grid, ((nv == 2 ? 4 : V.ncw) + 3) * 32, smem, ctx->stream>>>(P);
        NSK_CUDA(ctx, cudaGetLastError());
    }
    ctx->launches++;
    if (P.timing) {  // debugging aid: per-level averages of the stage cycle on stderr
        std::vector<unsigned long long> h((size_t)plan->grid * 16);
        NSK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        NSK_CUDA(ctx, cudaMemcpy(h.data(), P.timing, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int l = 0; l < k; l++) {
            double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double items = 0;
            double cyc_lo = 1e30, cyc_hi = 0.0, it_lo = 1e30, it_hi = 0.0;  // spread over the team's CTAs
            for (int b = 0; b < plan->grid; b++) {
                if ((int)h[(size_t)b * 16 + 9] != l || h[(size_t)b * 16 + 8] == 0) continue;
                for (int j = 0; j < 8; j++) sum[j] += (double)(long long)h[(size_t)b * 16 + j];
                const double nb = (double)h[(size_t)b * 16 + 8];
                items += nb;
                const double cyc = (double)(long long)h[(size_t)b * 16 + 4] / nb;
                cyc_lo = std::min(cyc_lo, cyc); cyc_hi = std::max(cyc_hi, cyc);
                it_lo = std::min(it_lo, nb); it_hi = std::max(it_hi, nb);
            }
            if (items > 0)
                fprintf(stderr, "pk_timing k=%d level %d: per-CTA mean cycle %.0f .. %.0f ns, items per CTA %.0f .. %.0f\n", k, l,
                        cyc_lo, cyc_hi, it_lo, it_hi);
            if (items > 0)
                fprintf(stderr, "pk_timing k=%d level %d team %d: per item ns: blob->x issue %.0f | x issue->full %.0f | consume %.0f | "
                        "w0 done->all done seen %.0f | cycle %.0f | fence+red(per item) %.0f | D wait stage %.0f | D wait deps %.0f\n",
                        k, l, plan->teams[l], sum[0] / items, sum[1] / items, sum[2] / items, sum[3] / items, sum[4] / items,
                        sum[5] / items, sum[6] / items, sum[7] / items);
        }
    }
    return NSK_OK;
}

bool nsk_packed_applicable(nsk_csr_t A)
{
    if (A->n == 0 || A->nnz == 0) return false;
    return pk_get(A, g_pkv[pk_variant(A)])->ok;
}

// bytes of the packed operator (for traffic accounting in the bench); 0 when not packed
size_t nsk_packed_bytes(nsk_csr_t A)
{
    PackedOp *op = pk_get(A, g_pkv[pk_variant(A)]);
    return op->ok ? op->blob_bytes : 0;
}
