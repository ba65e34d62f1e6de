// mpk_wavefront.cu -- fused matrix-powers kernel: levels[l] = A^(l+1) x for l = 0..k-1 in ONE
// persistent launch that reads the operator from HBM once.
//
// Replaces the reference's fused first-touch kernels SpM2V_CSR* (mpk/SpM2V.cpp:80-332), SpM3V and
// SpM4V (mpk/SpMVmulti0.cpp:132-221).  The reference keeps locality by a serial lazy traversal
// (compute row j of the lower level the first time some row i needs it); the B200 version keeps it
// by ORDER: work items are (level, tile) pairs issued along the skewed wavefront
//
//        tau = tile + level * D          (D = forward reach of the sparsity pattern, in tiles)
//
// so that power l of tile t runs shortly after power l-1 of the tiles it depends on, while their
// col/val slices (12 B/nnz, the dominant traffic) and the freshly written level vectors are still in
// the 126 MB L2.  HBM sees 12 nnz + 4(n+1) + 8n read once and 8nk written once (DESIGN.md section
// 4); the k-1 re-reads are served by L2 (measured 17-34 TB/s vs 7.3 TB/s HBM, profiles/).
//
// Mechanics
//   * same tiles, same shared-memory stage and same TMA bulk-copy ring as the streaming SpMV kernel;
//   * a CTA's producer lane claims the next work item with one atomicAdd on a global counter
//     (items are claimed strictly in wavefront order, which is what makes spinning deadlock-free:
//     the oldest unfinished item is always at the head of some CTA's ring and all its inputs are
//     older, hence finished) and prefetches its matrix slice immediately -- the slice does not
//     depend on any flag;
//   * a dedicated dependency warp acquires, for each claimed item (l, t), l >= 1, the completion counters
//     of the level l-1 tile GROUPS covering t's column range (ld.acquire.gpu, bounded spin) and then opens
//     the item for the consumer warps through an mbarrier -- the wait overlaps the previous item's work;
//   * consumer warps gather with ordinary coherent loads and, when their rows are stored, only arrive on
//     a CTA-local mbarrier; a publisher thread turns that into one red.release.gpu on the group counter
//     (the GPU-scope fence costs ~1 us -- measured 23 % of all stall samples when consumers paid it).
//   * per-row arithmetic is the same sequential chain as nsk_spmv, so every level is bit-identical
//     to k separate products (tests/test_spmv_gpu.py::test_mpk_wavefront_*).
#include <algorithm>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"
#include "stream_common.cuh"

using namespace nskptx;

std::vector<int> &nsk_csr_host_ptrow(nsk_csr_t A);
void nsk_pipe_free(nsk_csr_t A);  // mpk_pipeline.cu

#include "wave_common.h"

struct WaveTask {               // 32 bytes: everything an item needs in one go (two 16-byte loads)
    int row0, nrows, nz0, nz1;  // the tile (copied from the tiling so an item costs no second lookup)
    int level, tile, glo, ghi;  // inputs: groups [0, ghi] of level-1 must be complete; level < 0 = end marker
};

struct WaveParams {
    const WaveTask *tasks;  // [grid][per_cta]: CTA c runs tasks[c * per_cta + 0 ...] in order (static schedule)
    int per_cta;
    const WaveTask *tasks_dyn;  // items in wavefront order (dynamic schedule); ntasks of them
    int ntasks;
    unsigned int *next;         // dynamic schedule: claim cursor, zeroed before the launch; nullptr = static
    int ngroups;
    int *counters;          // [k][ngroups], zeroed before the launch
    const int *group_size;  // [k][ngroups] tiles that will report per group and level
    const int *ptrow;
    const int *indcol;
    const double *coef;
    const double *x;
    double *levels[NSK_MAX_K];
    int level_rows[NSK_MAX_K];
    int k;
};

template <int T_NNZ, int T_ROWS, int STAGES, int NCW, int MINB, int RPT, bool MULADD>
__global__ void __launch_bounds__((NCW + 3) * 32, MINB) mpk_wavefront_kernel(const WaveParams P)
{
    using Geo = StageGeom<T_NNZ, T_ROWS>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage_base = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Geo::BYTES * STAGES);  // TMA bytes of the stage landed
    uint64_t *empty = full + STAGES;                                            // all consumer warps are done with it
    uint64_t *claimed = empty + STAGES;                                         // header written (producer -> helpers)
    uint64_t *ready = claimed + STAGES;                                         // inputs of the item are complete
    uint64_t *done = ready + STAGES;                                            // all consumer warps stored their rows

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);      // publisher: stage header read, consumers finished
            mbar_init(&claimed[s], 1);
            mbar_init(&ready[s], 1);
            mbar_init(&done[s], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ===== producer.  The schedule is STATIC: this CTA's items sit in wavefront order in its own slice of
        // P.tasks, so nothing on this path waits for a global atomic (a dynamic claim = atomic + dependent
        // descriptor load capped every geometry at ~1 item/us per CTA, profiles/r01_sweep_mpk_wavefront_*).
        // Descriptors are fetched 32 at a time, one per lane, one batch ahead, and handed to lane 0 by shuffle;
        // an item's matrix slice is prefetched as soon as a stage is free -- it depends on no flag. =====
        if (P.next != nullptr) {
            // dynamic schedule: items are claimed in wavefront order with one atomicAdd each, which balances
            // load across SMs by itself.  A claim is two dependent L2 round trips (cursor, then descriptor):
            // claims are therefore PIPELINED three items ahead (cursor values q1..q3 and descriptor d0 are in
            // flight while the current item is issued), so the loop never waits for either.
            if (lane == 0) {
                const int4 *tasks4 = reinterpret_cast<const int4 *>(P.tasks_dyn);
                const unsigned int nt = (unsigned int)P.ntasks;
                const int4 endA = make_int4(0, 0, 0, 0), endB = make_int4(-1, 0, 0, 0);
                unsigned int q0 = atomicAdd(P.next, 1u);
                unsigned int q1 = atomicAdd(P.next, 1u);
                unsigned int q2 = atomicAdd(P.next, 1u);
                int4 ta = endA, tb = endB;
                if (q0 < nt) { ta = tasks4[2 * (size_t)q0]; tb = tasks4[2 * (size_t)q0 + 1]; }
                for (int it = 0;; ++it) {
                    const unsigned int q3 = atomicAdd(P.next, 1u);
                    int4 na = endA, nb = endB;
                    if (q1 < nt) { na = tasks4[2 * (size_t)q1]; nb = tasks4[2 * (size_t)q1 + 1]; }
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
                    int *hdr = reinterpret_cast<int *>(st + Geo::HDR_OFF);
                    hdr[0] = ta.x; hdr[1] = ta.y; hdr[2] = ta.z; hdr[3] = ta.w;
                    hdr[4] = tb.x; hdr[5] = tb.y; hdr[6] = tb.z; hdr[7] = tb.w;
                    mbar_arrive(&claimed[s]);
                    if (tb.x < 0) {  // end marker
                        mbar_arrive(&full[s]);
                        break;
                    }
                    const int row0 = ta.x, nrows = ta.y, nz0 = ta.z, nz1 = ta.w;
                    const int a0 = nz0 & ~3, v0 = nz0 & ~1, p0 = row0 & ~3;
                    const uint32_t cbytes = (uint32_t)(((nz1 - a0) + 3) & ~3) * 4u;
                    const uint32_t vbytes = (uint32_t)(((nz1 - v0) + 1) & ~1) * 8u;
                    const uint32_t pbytes = (uint32_t)(((row0 + nrows + 1 - p0) + 3) & ~3) * 4u;
                    mbar_arrive_expect_tx(&full[s], cbytes + vbytes + pbytes);
                    bulk_g2s(st + Geo::PTR_OFF, P.ptrow + p0, pbytes, &full[s]);
                    if (cbytes) bulk_g2s(st + Geo::COL_OFF, P.indcol + a0, cbytes, &full[s]);
                    if (vbytes) bulk_g2s(st + Geo::VAL_OFF, P.coef + v0, vbytes, &full[s]);
                    ta = na; tb = nb;
                    q1 = q2; q2 = q3;
                }
            }
            return;
        }
        const int4 *my = reinterpret_cast<const int4 *>(P.tasks) + 2 * (size_t)blockIdx.x * P.per_cta;
        const int4 endA = make_int4(0, 0, 0, 0), endB = make_int4(-1, 0, 0, 0);
        int4 ca = endA, cb = endB, na = endA, nb = endB;
        if (lane < P.per_cta) { ca = my[2 * lane]; cb = my[2 * lane + 1]; }
        if (32 + lane < P.per_cta) { na = my[2 * (32 + lane)]; nb = my[2 * (32 + lane) + 1]; }
        for (int it = 0;; ++it) {
            const int j = it & 31;
            if (j == 0 && it > 0) {
                ca = na; cb = nb;
                na = endA; nb = endB;
                const int idx = it + 32 + lane;
                if (idx < P.per_cta) { na = my[2 * idx]; nb = my[2 * idx + 1]; }
            }
            int4 ta, tb;
            ta.x = __shfl_sync(0xffffffffu, ca.x, j); ta.y = __shfl_sync(0xffffffffu, ca.y, j);
            ta.z = __shfl_sync(0xffffffffu, ca.z, j); ta.w = __shfl_sync(0xffffffffu, ca.w, j);
            tb.x = __shfl_sync(0xffffffffu, cb.x, j); tb.y = __shfl_sync(0xffffffffu, cb.y, j);
            tb.z = __shfl_sync(0xffffffffu, cb.z, j); tb.w = __shfl_sync(0xffffffffu, cb.w, j);
            if (lane == 0) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
                int *hdr = reinterpret_cast<int *>(st + Geo::HDR_OFF);
                hdr[0] = ta.x; hdr[1] = ta.y; hdr[2] = ta.z; hdr[3] = ta.w;
                hdr[4] = tb.x; hdr[5] = tb.y; hdr[6] = tb.z; hdr[7] = tb.w;
                mbar_arrive(&claimed[s]);
                if (tb.x < 0) {  // end marker
                    mbar_arrive(&full[s]);
                } else {
                    const int row0 = ta.x, nrows = ta.y, nz0 = ta.z, nz1 = ta.w;
                    const int a0 = nz0 & ~3, v0 = nz0 & ~1, p0 = row0 & ~3;
                    const uint32_t cbytes = (uint32_t)(((nz1 - a0) + 3) & ~3) * 4u;
                    const uint32_t vbytes = (uint32_t)(((nz1 - v0) + 1) & ~1) * 8u;
                    const uint32_t pbytes = (uint32_t)(((row0 + nrows + 1 - p0) + 3) & ~3) * 4u;
                    mbar_arrive_expect_tx(&full[s], cbytes + vbytes + pbytes);
                    bulk_g2s(st + Geo::PTR_OFF, P.ptrow + p0, pbytes, &full[s]);
                    if (cbytes) bulk_g2s(st + Geo::COL_OFF, P.indcol + a0, cbytes, &full[s]);
                    if (vbytes) bulk_g2s(st + Geo::VAL_OFF, P.coef + v0, vbytes, &full[s]);
                }
            }
            if (tb.x < 0) break;  // warp-uniform
        }
        return;
    }

    if (warp == NCW + 1) {
        // ===== dependency warp.  Lane l keeps a WATERMARK for level l: every tile group below it is known to
        // be complete.  An item whose last input group lies below the watermark opens at once (no memory
        // traffic); otherwise the warp polls the next 32 group counters in one round trip (ld.acquire.gpu,
        // bounded spin) and advances the watermark over the complete prefix -- the lower level normally runs
        // `slack` tiles ahead, so one poll opens several items.  Items of one CTA are in wavefront order and all
        // inputs of the oldest unfinished item are older, hence finished: spinning cannot deadlock. =====
        int W = 0;
        for (int it = 0;; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&claimed[s], ph);
            const int *hdr = reinterpret_cast<const int *>(stage_base + (size_t)s * Geo::BYTES + Geo::HDR_OFF);
            const int level = hdr[4];
            if (level > 0) {
                const int ghi = hdr[7];
                int w = __shfl_sync(0xffffffffu, W, level - 1);
                if (w <= ghi) {
                    const int *cnt = P.counters + (size_t)(level - 1) * P.ngroups;
                    const int *need = P.group_size + (size_t)(level - 1) * P.ngroups;
                    uint32_t spins = 0;
                    while (true) {
                        const int g = w + lane;
                        bool ok = true;
                        // ld.acquire.gpu = LDG.STRONG.GPU + CCTL.IVALL: also drops this SM's stale L1 lines
                        if (g < P.ngroups) ok = ld_acquire_gpu(cnt + g) >= __ldg(need + g);
                        const unsigned int bad = __ballot_sync(0xffffffffu, !ok);
                        const int adv = bad ? __ffs(bad) - 1 : 32;
                        w = min(w + adv, P.ngroups);
                        if (w > ghi) break;
                        if (adv == 0) {
                            __nanosleep(64);
                            if (++spins > (1u << 24)) __trap();  // protocol bug: fail the launch, never hang
                        }
                    }
                    if (lane == level - 1) W = w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);  // release.cta: orders the acquires above before the consumers
            if (level < 0) break;
        }
        return;
    }

    if (warp == NCW + 2) {
        // ===== publisher: makes finished items visible to the rest of the GPU.  The GPU-scope fence (~1 us) is
        // paid here, off the consumers' path, and ONCE for all items found finished at that moment: consumers
        // only arrive on done[s] (release.cta); this thread's acquire of done[s] followed by fence + RED is
        // cumulative over their stores. =====
        if (lane == 0) {
            for (int it = 0;;) {
                int slot[STAGES];  // counter index + 1, 0 = nothing to publish (last level)
                int n = 0;
                bool end = false;
#pragma unroll
                for (int j = 0; j < STAGES; j++) {
                    if (j > n || end) continue;  // stop at the first item that is not finished yet
                    const int s = (it + j) % STAGES;
                    const uint32_t ph = ((it + j) / STAGES) & 1;
                    if (j == 0) mbar_wait(&claimed[s], ph);
                    else if (!mbar_try_wait(&claimed[s], ph)) continue;
                    const int *hdr = reinterpret_cast<const int *>(stage_base + (size_t)s * Geo::BYTES + Geo::HDR_OFF);
                    const int level = hdr[4], tile = hdr[5];
                    if (level < 0) { end = true; continue; }
                    if (j == 0) mbar_wait(&done[s], ph);
                    else if (!mbar_try_wait(&done[s], ph)) continue;
                    slot[j] = level < P.k - 1 ? level * P.ngroups + tile / WF_GROUP + 1 : 0;
                    mbar_arrive(&empty[s]);  // the producer may refill the stage while we publish
                    n = j + 1;
                }
                if (n == 0) break;  // first item was the end marker
                __threadfence();
#pragma unroll
                for (int j = 0; j < STAGES; j++)
                    if (j < n && slot[j]) red_relaxed_gpu_add(P.counters + (slot[j] - 1), 1);
                it += n;
            }
        }
        return;
    }

    // ===== consumer warps: independent of one another (no CTA-wide barrier in the loop) =====
    constexpr int NCT = NCW * 32;
    for (int it = 0;; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&ready[s], ph);
        mbar_wait(&full[s], ph);
        unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
        const int *hdr = reinterpret_cast<const int *>(st + Geo::HDR_OFF);
        const int level = hdr[4];
        if (level < 0) break;
        const int row0 = hdr[0], nrows = hdr[1], nz0 = hdr[2];
        const double *val_s = reinterpret_cast<const double *>(st + Geo::VAL_OFF);
        const int *col_s = reinterpret_cast<const int *>(st + Geo::COL_OFF);
        const int *ptr_s = reinterpret_cast<const int *>(st + Geo::PTR_OFF);
        const int vo = nz0 & ~1, co = nz0 & ~3, po = row0 & ~3;
        // level 0 reads x (constant for the launch: read-only path); level l >= 1 reads what other CTAs wrote
        // earlier in THIS launch: ordinary coherent loads, ordered by ready[s] after the dependency warp's
        // acquires (which also invalidated this SM's L1)
        if (level == 0)
            consume_tile<MULADD, true, RPT, NCT>(val_s, col_s, ptr_s, vo, co, po, row0, nrows, P.level_rows[0], P.x,
                                                 P.levels[0], tid);
        else
            consume_tile<MULADD, false, RPT, NCT>(val_s, col_s, ptr_s, vo, co, po, row0, nrows, P.level_rows[level],
                                                  P.levels[level - 1], P.levels[level], tid);
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);  // release.cta; no GPU-scope fence on this path
    }
}

// -----------------------------------------------------------------------------------------------
// host side: plan (cached per operator / k / level_rows) and launch
// -----------------------------------------------------------------------------------------------
struct WavePlan {
    int k = 0;
    int t_nnz = 0, t_rows = 0;
    int slack = 0;
    int grid_req = 0;  // resident CTAs the plan was asked for (cache key)
    int l2_pct = 0;    // L2 budget option the plan was built / refused under (cache key)
    int grid = 0;      // CTAs it uses: min(grid_req, items)
    int per_cta = 0;
    bool rejected = false;
    std::vector<int> level_rows;
    int ntasks = 0, ngroups = 0, D = 0;
    WaveTask *d_tasks = nullptr;
    WaveTask *d_tasks_dyn = nullptr;
    int *d_counters = nullptr;    // k*ngroups ints + 1 cursor (last)
    int *d_group_size = nullptr;
};

struct WaveState {
    std::vector<WavePlan> plans;
    // column extents per row block (<= 32 rows, never straddling a break), in GLOBAL-ORDER rank space
    std::vector<int> blk_row0, blk_min, blk_max;
};

#include <map>
#include <mutex>
static std::map<nsk_csr_t, WaveState> g_wave;  // map guarded; an entry belongs to its operator's thread
static std::mutex g_wave_mu;
static WaveState &wave_state(nsk_csr_t A)
{
    std::lock_guard<std::mutex> lk(g_wave_mu);
    return g_wave[A];
}

// Called at create time (and again by the distributed layer once breaks / row_rank are known).
// Columns >= A->n are ghost entries of x that only level 0 reads: they create no dependency.
void nsk_wave_set_block_extents(nsk_csr_t A, const int *ptrow, const int *indcol)
{
    WaveState &S = wave_state(A);
    nsk_pipe_free(A);  // plans of the level-pipeline kernel were built from the old extents
    nsk_packed_free(A);
    nsk_sell_free(A);
    S.plans.clear();
    S.blk_row0.clear(); S.blk_min.clear(); S.blk_max.clear();
    const int n = A->n;
    const bool ranked = !A->row_rank.empty();
    size_t bi = 0;
    int r = 0;
    while (r < n) {
        while (bi < A->breaks.size() && A->breaks[bi] <= r) bi++;
        const int seg_end = bi < A->breaks.size() ? std::min(n, A->breaks[bi]) : n;
        const int r1 = std::min(seg_end, r + 32);
        int mn = INT32_MAX, mx = -1;
        for (int j = ptrow[r]; j < ptrow[r1]; j++) {
            int c = indcol[j];
            if (c >= n) continue;
            if (ranked) c = A->row_rank[c];
            mn = c < mn ? c : mn;
            mx = c > mx ? c : mx;
        }
        S.blk_row0.push_back(r);
        S.blk_min.push_back(mn);
        S.blk_max.push_back(mx);
        r = r1;
    }
}

void nsk_wave_free(nsk_csr_t A)
{
    std::vector<WavePlan> mine;
    {
        std::lock_guard<std::mutex> lk(g_wave_mu);
        auto it = g_wave.find(A);
        if (it == g_wave.end()) return;
        mine.swap(it->second.plans);
        g_wave.erase(it);
    }
    for (WavePlan &p : mine) {
        if (p.d_tasks) cudaFree(p.d_tasks);
        if (p.d_tasks_dyn) cudaFree(p.d_tasks_dyn);
        if (p.d_counters) cudaFree(p.d_counters);
        if (p.d_group_size) cudaFree(p.d_group_size);
    }
}

// Dependency geometry of a tiling, shared by the wavefront and the level-pipeline kernels: tile order in
// global row order (positions), and for every tile the range of position GROUPS its columns fall into.
bool nsk_wave_deps(nsk_csr_t A, const nsk_tiling &T, WaveDeps &out, const char **why)
{
    WaveState &S = wave_state(A);
    if (S.blk_row0.empty()) { *why = "column extents were not recorded"; return false; }
    if (T.nlong) { *why = "operator has rows longer than a stage"; return false; }
    const int ntiles = T.ntiles;
    if (ntiles == 0) { *why = "empty operator"; return false; }
    const int ngroups = (ntiles + WF_GROUP - 1) / WF_GROUP;
    const bool ranked = !A->row_rank.empty();

    // POSITION of a tile = its place in global row order (identity for a single-GPU operator; for a
    // distributed slab the ghost rings, stored after the owned rows, slot in below / above them).
    // Tiles never straddle a break, so a tile is a contiguous run in rank space too.
    std::vector<int> tile_at_pos(ntiles), pos_of_tile(ntiles), key(ntiles);
    for (int t = 0; t < ntiles; t++) {
        tile_at_pos[t] = t;
        key[t] = ranked ? A->row_rank[T.h_tiles[t].row0] : T.h_tiles[t].row0;
    }
    if (ranked) std::sort(tile_at_pos.begin(), tile_at_pos.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<int> pos_key(ntiles);
    for (int p = 0; p < ntiles; p++) {
        pos_of_tile[tile_at_pos[p]] = p;
        pos_key[p] = key[tile_at_pos[p]];
    }
    auto pos_of_rank = [&](int rank) {
        int p = (int)(std::upper_bound(pos_key.begin(), pos_key.end(), rank) - pos_key.begin()) - 1;
        return p < 0 ? 0 : p;
    };
    std::vector<int> glo(ntiles), ghi(ntiles);  // indexed by tile
    int reach = 0;
    for (int t = 0; t < ntiles; t++) {
        const nsk_tile &tl = T.h_tiles[t];
        int mn = INT32_MAX, mx = -1;
        size_t b = (size_t)(std::upper_bound(S.blk_row0.begin(), S.blk_row0.end(), tl.row0) - S.blk_row0.begin()) - 1;
        for (; b < S.blk_row0.size() && S.blk_row0[b] < tl.row0 + tl.nrows; b++) {
            mn = std::min(mn, S.blk_min[b]);
            mx = std::max(mx, S.blk_max[b]);
        }
        if (mx < 0) { mn = key[t]; mx = key[t]; }  // rows without (local) entries depend on nothing
        glo[t] = pos_of_rank(mn) / WF_GROUP;
        ghi[t] = pos_of_rank(mx) / WF_GROUP;
        const int last_needed = std::min(ntiles - 1, (ghi[t] + 1) * WF_GROUP - 1);
        reach = std::max(reach, last_needed - pos_of_tile[t]);
    }
    out.ntiles = ntiles;
    out.ngroups = ngroups;
    out.reach = reach;
    out.tile_at_pos.swap(tile_at_pos);
    out.pos_of_tile.swap(pos_of_tile);
    out.glo.swap(glo);
    out.ghi.swap(ghi);
    return true;
}

struct WaveVariant {
    int t_nnz, t_rows, stages, ncw, minb, rpt;
};
//                 T_NNZ T_ROWS STAGES NCW MINB RPT(rows per consumer thread and pass)
#define NSK_WAVE_VARIANTS(X) \
    X(0, 2048, 256, 3, 8, 2, 1)   \
    X(1, 2048, 256, 4, 8, 2, 1)   \
    X(2, 4096, 512, 4, 16, 1, 1)  \
    X(3, 4096, 512, 4, 8, 1, 2)   \
    X(4, 3584, 512, 5, 16, 1, 1)  \
    X(5, 3584, 512, 5, 8, 1, 2)   \
    X(6, 7168, 1024, 2, 16, 1, 2) \
    X(7, 1792, 256, 3, 8, 3, 1)   \
    X(8, 1792, 256, 6, 8, 2, 1)   \
    X(9, 1792, 256, 10, 8, 1, 1)  \
    X(10, 1024, 128, 4, 4, 4, 1)  \
    X(11, 3584, 512, 5, 28, 1, 1)

static const WaveVariant g_wvariants[] = {
#define X(id, t, r, s, w, b, u) {t, r, s, w, b, u},
    NSK_WAVE_VARIANTS(X)
#undef X
};
static const int g_nwvariants = sizeof(g_wvariants) / sizeof(g_wvariants[0]);

typedef void (*wave_fn)(const WaveParams);
static wave_fn wave_lookup(int variant, bool muladd, int *smem)
{
    switch (variant) {
#define X(id, t, r, s, w, b, u)                                                    \
    case id:                                                                       \
        *smem = StageGeom<t, r>::BYTES * s + 5 * s * 8 + 128;                      \
        return muladd ? mpk_wavefront_kernel<t, r, s, w, b, u, true> : mpk_wavefront_kernel<t, r, s, w, b, u, false>;
        NSK_WAVE_VARIANTS(X)
#undef X
    }
    return nullptr;
}


static int wave_variant(nsk_ctx_t ctx)
{
    // option value 0 = default; n >= 1 selects table entry n - 1
    int v = (int)ctx->opt.wave_variant - 1;
    if (v < 0 || v >= g_nwvariants) v = 2;
    return v;
}

// Builds (or finds) the plan.  Returns nullptr (and leaves *why) when the wavefront does not apply.
// slack: extra skew (in tiles) between consecutive levels on top of the pattern's reach.  Work items are
// claimed in wavefront order but complete a few microseconds later (grid x STAGES items are in flight);
// without slack every level-l item would wait for a level-(l-1) item claimed one step earlier and the
// whole sweep would serialise on that latency (measured: 16 ms instead of 0.5 ms on 256^3, k = 4).
static WavePlan *get_plan(nsk_csr_t A, int k, const int *level_rows, const WaveVariant &V, int slack, int grid,
                          const char **why)
{
    WaveState &S = wave_state(A);
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : A->n;
    for (WavePlan &p : S.plans)
        if (p.k == k && p.t_nnz == V.t_nnz && p.t_rows == V.t_rows && p.slack == slack && p.grid_req == grid && p.level_rows == lr &&
            p.l2_pct == (int)A->ctx->opt.wave_l2_pct) {
            if (p.rejected) { *why = "wavefront window exceeds the L2 budget"; return nullptr; }
            return &p;
        }

    if (S.blk_row0.empty()) { *why = "column extents were not recorded"; return nullptr; }
    const nsk_tiling *Tp = nullptr;
    if (nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp) != NSK_OK) { *why = "tiling failed"; return nullptr; }
    const nsk_tiling &T = *Tp;
    WaveDeps Dp;
    if (!nsk_wave_deps(A, T, Dp, why)) return nullptr;
    const int ntiles = Dp.ntiles, ngroups = Dp.ngroups, reach = Dp.reach;
    const std::vector<int> &tile_at_pos = Dp.tile_at_pos, &glo = Dp.glo, &ghi = Dp.ghi;
    const int D = reach + 1 + slack;
    // The window that must stay in L2: (k-1)*D tiles of matrix data plus k level vectors of it.
    const double tile_bytes = 12.0 * A->mean_row * V.t_rows + 8.0 * V.t_rows * (k + 1);
    const double window = (double)(k - 1) * D * tile_bytes;
    const double budget = (A->ctx->opt.wave_l2_pct > 0 ? (double)A->ctx->opt.wave_l2_pct : 80.0) / 100.0;
    if (window > budget * (double)A->ctx->prop.l2CacheSize) {
        *why = "wavefront window exceeds the L2 budget";
        WavePlan rej;  // remember the refusal: planning costs O(tiles) host work
        rej.k = k; rej.t_nnz = V.t_nnz; rej.t_rows = V.t_rows; rej.level_rows = lr; rej.slack = slack; rej.grid_req = grid; rej.rejected = true;
        rej.l2_pct = (int)A->ctx->opt.wave_l2_pct;
        S.plans.push_back(rej);
        return nullptr;
    }

    std::vector<WaveTask> tasks;
    tasks.reserve((size_t)k * ntiles);
    std::vector<int> gsize((size_t)k * ngroups, 0);
    const int tau_end = ntiles + (k - 1) * D;
    for (int tau = 0; tau < tau_end; tau++)
        for (int l = k - 1; l >= 0; l--) {
            const int p = tau - l * D;
            if (p < 0 || p >= ntiles) continue;
            const int t = tile_at_pos[p];
            const nsk_tile &tl = T.h_tiles[t];
            if (tl.row0 >= lr[l]) continue;  // outside this level's row prefix (distributed shrink)
            tasks.push_back(WaveTask{tl.row0, tl.nrows, tl.nz0, tl.nz1, l, p, glo[t], ghi[t]});
            gsize[(size_t)l * ngroups + p / WF_GROUP]++;
        }

    // Static schedule: item q of the wavefront order belongs to CTA (q + q / G) mod G, G = min(grid, items) -- round
    // robin, rotated by one every round so that no CTA is pinned to one level (with G a multiple of k a plain
    // round robin would leave the HBM-fed level 0 to G/k SMs).  Each CTA's items are stored contiguously, in
    // order, closed by an end marker.
    const int ntasks = (int)tasks.size();
    const int G = std::max(1, std::min(grid, ntasks));
    const int rounds = (ntasks + G - 1) / G;
    const int per_cta = rounds + 1;
    std::vector<WaveTask> sched((size_t)G * per_cta, WaveTask{0, 0, 0, 0, -1, 0, 0, 0});
    for (int q = 0; q < ntasks; q++) {
        const int r = q / G;
        const int c = (q + r) % G;
        sched[(size_t)c * per_cta + r] = tasks[q];
    }

    WavePlan p;
    p.k = k; p.t_nnz = V.t_nnz; p.t_rows = V.t_rows; p.level_rows = lr; p.slack = slack; p.grid_req = grid; p.grid = G; p.per_cta = per_cta;
    p.l2_pct = (int)A->ctx->opt.wave_l2_pct;
    p.ntasks = ntasks; p.ngroups = ngroups; p.D = D;
    if (cudaMalloc(&p.d_tasks, sizeof(WaveTask) * sched.size()) != cudaSuccess ||
        cudaMalloc(&p.d_tasks_dyn, sizeof(WaveTask) * (tasks.size() + 1)) != cudaSuccess ||
        cudaMalloc(&p.d_counters, sizeof(int) * ((size_t)k * ngroups + 4)) != cudaSuccess ||
        cudaMalloc(&p.d_group_size, sizeof(int) * (size_t)k * ngroups) != cudaSuccess) {
        *why = "plan allocation failed";
        return nullptr;
    }
    cudaMemcpy(p.d_tasks, sched.data(), sizeof(WaveTask) * sched.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_tasks_dyn, tasks.data(), sizeof(WaveTask) * tasks.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_group_size, gsize.data(), sizeof(int) * gsize.size(), cudaMemcpyHostToDevice);
    S.plans.push_back(p);
    return &S.plans.back();
}

// resident CTAs of the chosen kernel and the slack (tiles) that keeps dependent levels apart
static int wave_launch_shape(nsk_ctx_t ctx, int variant, bool muladd, int k, wave_fn *fn_out, int *smem_out, int *grid_max,
                             int *slack)
{
    const WaveVariant &V = g_wvariants[variant];
    int smem = 0;
    wave_fn fn = wave_lookup(variant, muladd, &smem);
    NSK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    NSK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, (V.ncw + 3) * 32, smem));
    if (ctx->opt.spmv_ctas_per_sm > 0) per_sm = std::min(per_sm, (int)ctx->opt.spmv_ctas_per_sm);
    NSK_REQUIRE(ctx, per_sm >= 1, "wavefront kernel does not fit on an SM");
    *grid_max = ctx->prop.multiProcessorCount * per_sm;
    const double pct = ctx->opt.wave_slack_pct >= 0 ? (double)ctx->opt.wave_slack_pct : 100.0;
    *slack = (int)((pct / 100.0) * (double)(*grid_max) * (V.stages + (ctx->opt.wave_static ? 0 : 3)) / (double)k + 0.999);  // dynamic: + the 3 pre-claimed items
    *fn_out = fn;
    *smem_out = smem;
    return NSK_OK;
}

bool nsk_mpk_wavefront_applicable(nsk_csr_t A, int k)
{
    if (k < 2 || A->mean_row > 64.0 || A->n == 0) return false;
    const char *why = nullptr;
    wave_fn fn; int smem, grid_max, slack;
    const int variant = wave_variant(A->ctx);
    if (wave_launch_shape(A->ctx, variant, false, k, &fn, &smem, &grid_max, &slack) != NSK_OK) return false;
    return get_plan(A, k, nullptr, g_wvariants[variant], slack, grid_max, &why) != nullptr;
}

int nsk_mpk_wavefront(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                      const int *level_rows)
{
    nsk_ctx_t ctx = A->ctx;
    const int variant = wave_variant(ctx);
    const WaveVariant &V = g_wvariants[variant];
    const char *why = "";
    wave_fn fn; int smem = 0, grid_max = 0, slack = 0;
    NSK_TRY(wave_launch_shape(ctx, variant, mode == NSK_EXACT_MULADD, k, &fn, &smem, &grid_max, &slack));
    WavePlan *plan = get_plan(A, k, level_rows, V, slack, grid_max, &why);
    if (!plan) {
        nsk_set_error(ctx, "wavefront matrix powers not applicable: %s", why);
        return NSK_ERR_UNSUPPORTED;
    }
    const nsk_tiling *Tp = nullptr;
    NSK_TRY(nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp));
    const int grid = plan->grid;

    const size_t ncnt = (size_t)k * plan->ngroups;
    NSK_CUDA(ctx, cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (ncnt + 4), ctx->stream));
    WaveParams P;
    P.tasks = plan->d_tasks;
    P.per_cta = plan->per_cta;
    P.tasks_dyn = plan->d_tasks_dyn;
    P.ntasks = plan->ntasks;
    P.next = ctx->opt.wave_static ? nullptr : reinterpret_cast<unsigned int *>(plan->d_counters + ncnt);
    P.ngroups = plan->ngroups;
    P.counters = plan->d_counters;
    P.group_size = plan->d_group_size;
    P.ptrow = A->d_ptrow;
    P.indcol = A->d_indcol;
    P.coef = A->d_coef;
    P.x = d_x;
    for (int l = 0; l < NSK_MAX_K; l++) {
        P.levels[l] = l < k ? d_levels[l] : nullptr;
        P.level_rows[l] = l < k ? plan->level_rows[l] : 0;
    }
    P.k = k;
    fn<<<grid, (V.ncw + 3) * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}
