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

constexpr int WF_GROUP = 16;  // tiles per completion counter

struct WaveTask {               // 32 bytes: everything a claim needs in one go (two 16-byte loads)
    int row0, nrows, nz0, nz1;  // the tile (copied from the tiling so a claim costs no second lookup)
    int level, tile, glo, ghi;  // inputs: groups [glo, ghi] of level-1 must be complete
};

struct WaveParams {
    const nsk_tile *tiles;
    const WaveTask *tasks;
    int ntasks;
    int ngroups;
    int *counters;          // [k][ngroups], zeroed before the launch
    const int *group_size;  // [k][ngroups] tiles that will report per group and level
    unsigned int *next;     // work-item cursor, zeroed before the launch
    const int *ptrow;
    const int *indcol;
    const double *coef;
    const double *x;
    double *levels[NSK_MAX_K];
    int level_rows[NSK_MAX_K];
    int k;
};

template <int T_NNZ, int T_ROWS, int STAGES, int NCW, int MINB, bool MULADD>
__global__ void __launch_bounds__((NCW + 3) * 32, MINB) mpk_wavefront_kernel(const WaveParams P)
{
    using Geo = StageGeom<T_NNZ, T_ROWS>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage_base = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Geo::BYTES * STAGES);  // TMA bytes of the stage landed
    uint64_t *empty = full + STAGES;                                            // all consumer warps are done with it
    uint64_t *claimed = empty + STAGES;                                         // header written (producer -> dependency warp)
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
        // ===== producer: claims items in wavefront order, prefetches their matrix slices.  The claim
        // (one L2 atomic + one dependent 32-byte load, ~1.5 us under load) is taken ONE ITEM AHEAD, in the
        // shadow of the wait for a free stage: with the claim inside the stage turn-around the whole
        // sweep ran at ~0.5 item/us per CTA whatever the geometry (measured, profiles/). =====
        if (lane == 0) {
            const int4 *tasks4 = reinterpret_cast<const int4 *>(P.tasks);
            unsigned int q = atomicAdd(P.next, 1u);
            int4 ta = make_int4(0, 0, 0, 0), tb = make_int4(-1, 0, 0, 0);
            if (q < (unsigned int)P.ntasks) { ta = tasks4[2 * (size_t)q]; tb = tasks4[2 * (size_t)q + 1]; }
            for (int it = 0;; ++it) {
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
                const uint32_t cb = (uint32_t)(((nz1 - a0) + 3) & ~3) * 4u;
                const uint32_t vb = (uint32_t)(((nz1 - v0) + 1) & ~1) * 8u;
                const uint32_t pb = (uint32_t)(((row0 + nrows + 1 - p0) + 3) & ~3) * 4u;
                mbar_arrive_expect_tx(&full[s], cb + vb + pb);
                bulk_g2s(st + Geo::PTR_OFF, P.ptrow + p0, pb, &full[s]);
                if (cb) bulk_g2s(st + Geo::COL_OFF, P.indcol + a0, cb, &full[s]);
                if (vb) bulk_g2s(st + Geo::VAL_OFF, P.coef + v0, vb, &full[s]);
                // next claim, overlapped with the wait above on the next trip
                q = atomicAdd(P.next, 1u);
                tb.x = -1;
                if (q < (unsigned int)P.ntasks) { ta = tasks4[2 * (size_t)q]; tb = tasks4[2 * (size_t)q + 1]; }
            }
        }
        return;
    }

    if (warp == NCW + 1) {
        // ===== dependency warp: waits (off the consumers' critical path) until the level l-1 tile groups an
        // item reads from are complete, then opens the item for the consumers =====
        for (int it = 0;; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&claimed[s], ph);
            const int *hdr = reinterpret_cast<const int *>(stage_base + (size_t)s * Geo::BYTES + Geo::HDR_OFF);
            const int level = hdr[4];
            if (level > 0) {
                const int glo = hdr[6], ghi = hdr[7];
                const int *cnt = P.counters + (size_t)(level - 1) * P.ngroups;
                const int *need = P.group_size + (size_t)(level - 1) * P.ngroups;
                for (int g = glo + lane; g <= ghi; g += 32) {
                    const int want = need[g];
                    uint32_t spins = 0;
                    // ld.acquire.gpu = LDG.STRONG.GPU + CCTL.IVALL: also drops this SM's stale L1 lines
                    while (ld_acquire_gpu(cnt + g) < want) {
                        __nanosleep(32);
                        if (++spins > (1u << 24)) __trap();  // protocol bug: fail the launch, never hang
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);  // release.cta: orders the acquires above before the consumers
            if (level < 0) break;
        }
        return;
    }

    if (warp == NCW + 2) {
        // ===== publisher: makes a finished item visible to the rest of the GPU.  The GPU-scope release
        // (fence + RED, ~1 us) is paid here, off the consumers' path: consumers only arrive on done[s]
        // (release.cta); this thread's acquire of done[s] followed by red.release.gpu is cumulative over
        // their stores. =====
        if (lane == 0) {
            for (int it = 0;; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&claimed[s], ph);
                const int *hdr = reinterpret_cast<const int *>(stage_base + (size_t)s * Geo::BYTES + Geo::HDR_OFF);
                const int level = hdr[4], tile = hdr[5];
                if (level < 0) break;
                mbar_wait(&done[s], ph);
                mbar_arrive(&empty[s]);  // the producer may refill the stage while we publish
                if (level < P.k - 1) red_release_gpu_add(P.counters + (size_t)level * P.ngroups + tile / WF_GROUP, 1);
            }
        }
        return;
    }

    // ===== consumer warps: independent of one another (no CTA-wide barrier in the loop) =====
    constexpr int NCT = NCW * 32;
    const int ctid = tid;
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
        const int row_end = P.level_rows[level];
        double *dst = P.levels[level];
        // level 0 reads x (constant for the launch: read-only path); level l >= 1 reads what other CTAs wrote
        // earlier in THIS launch: ordinary coherent loads, ordered by ready[s] after the dependency warp's
        // acquires (which also invalidated this SM's L1)
        const double *src = level == 0 ? P.x : P.levels[level - 1];
        for (int r = ctid; r < nrows; r += NCT) {
            const int row = row0 + r;
            if (row >= row_end) continue;
            const int p = ptr_s[row - po], q = ptr_s[row + 1 - po];
            dst[row] = level == 0 ? row_chain<MULADD, true>(val_s, col_s, p, q, vo, co, src)
                                  : row_chain<MULADD, false>(val_s, col_s, p, q, vo, co, src);
        }
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
    bool rejected = false;
    std::vector<int> level_rows;
    int ntasks = 0, ngroups = 0, D = 0;
    WaveTask *d_tasks = nullptr;
    int *d_counters = nullptr;    // k*ngroups ints + 1 cursor (last)
    int *d_group_size = nullptr;
};

struct WaveState {
    std::vector<WavePlan> plans;
    // column extents per row block (<= 32 rows, never straddling a break), in GLOBAL-ORDER rank space
    std::vector<int> blk_row0, blk_min, blk_max;
};

#include <map>
static std::map<nsk_csr_t, WaveState> g_wave;

// Called at create time (and again by the distributed layer once breaks / row_rank are known).
// Columns >= A->n are ghost entries of x that only level 0 reads: they create no dependency.
void nsk_wave_set_block_extents(nsk_csr_t A, const int *ptrow, const int *indcol)
{
    WaveState &S = g_wave[A];
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
    auto it = g_wave.find(A);
    if (it == g_wave.end()) return;
    for (WavePlan &p : it->second.plans) {
        if (p.d_tasks) cudaFree(p.d_tasks);
        if (p.d_counters) cudaFree(p.d_counters);
        if (p.d_group_size) cudaFree(p.d_group_size);
    }
    g_wave.erase(it);
}

struct WaveVariant {
    int t_nnz, t_rows, stages, ncw, minb;
};
//                 T_NNZ T_ROWS STAGES NCW MINB
#define NSK_WAVE_VARIANTS(X) \
    X(0, 2048, 256, 2, 8, 3)   \
    X(1, 2048, 256, 3, 8, 2)   \
    X(2, 4096, 512, 2, 16, 2)  \
    X(3, 2048, 256, 2, 8, 4)   \
    X(4, 1024, 128, 3, 4, 5)   \
    X(5, 1024, 128, 2, 4, 6)   \
    X(6, 4096, 512, 3, 16, 1)  \
    X(7, 4096, 512, 4, 16, 1)  \
    X(8, 3584, 512, 5, 16, 1)  \
    X(9, 2048, 256, 4, 8, 2)   \
    X(10, 3072, 384, 3, 12, 1) \
    X(11, 3072, 384, 5, 12, 1)

static const WaveVariant g_wvariants[] = {
#define X(id, t, r, s, w, b) {t, r, s, w, b},
    NSK_WAVE_VARIANTS(X)
#undef X
};
static const int g_nwvariants = sizeof(g_wvariants) / sizeof(g_wvariants[0]);

typedef void (*wave_fn)(const WaveParams);
static wave_fn wave_lookup(int variant, bool muladd, int *smem)
{
    switch (variant) {
#define X(id, t, r, s, w, b)                                                       \
    case id:                                                                       \
        *smem = StageGeom<t, r>::BYTES * s + 5 * s * 8 + 128;                      \
        return muladd ? mpk_wavefront_kernel<t, r, s, w, b, true> : mpk_wavefront_kernel<t, r, s, w, b, false>;
        NSK_WAVE_VARIANTS(X)
#undef X
    }
    return nullptr;
}


static int wave_variant(nsk_ctx_t ctx)
{
    // option value 0 = default; n >= 1 selects table entry n - 1
    int v = (int)ctx->opt.wave_variant - 1;
    if (v < 0 || v >= g_nwvariants) v = 6;  // measured best on 256^3 (profiles/r01_sweep_mpk_wavefront_claimahead.txt)
    return v;
}

// Builds (or finds) the plan.  Returns nullptr (and leaves *why) when the wavefront does not apply.
// slack: extra skew (in tiles) between consecutive levels on top of the pattern's reach.  Work items are
// claimed in wavefront order but complete a few microseconds later (grid x STAGES items are in flight);
// without slack every level-l item would wait for a level-(l-1) item claimed one step earlier and the
// whole sweep would serialise on that latency (measured: 16 ms instead of 0.5 ms on 256^3, k = 4).
static WavePlan *get_plan(nsk_csr_t A, int k, const int *level_rows, const WaveVariant &V, int slack, const char **why)
{
    WaveState &S = g_wave[A];
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : A->n;
    for (WavePlan &p : S.plans)
        if (p.k == k && p.t_nnz == V.t_nnz && p.t_rows == V.t_rows && p.slack == slack && p.level_rows == lr) {
            if (p.rejected) { *why = "wavefront window exceeds the L2 budget"; return nullptr; }
            return &p;
        }

    if (S.blk_row0.empty()) { *why = "column extents were not recorded"; return nullptr; }
    const nsk_tiling *Tp = nullptr;
    if (nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp) != NSK_OK) { *why = "tiling failed"; return nullptr; }
    const nsk_tiling &T = *Tp;
    if (T.nlong) { *why = "operator has rows longer than a stage"; return nullptr; }
    const int ntiles = T.ntiles;
    if (ntiles == 0) { *why = "empty operator"; return nullptr; }
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
    const int D = reach + 1 + slack;
    // The window that must stay in L2: (k-1)*D tiles of matrix data plus k level vectors of it.
    const double tile_bytes = 12.0 * A->mean_row * V.t_rows + 8.0 * V.t_rows * (k + 1);
    const double window = (double)(k - 1) * D * tile_bytes;
    const double budget = (A->ctx->opt.wave_l2_pct > 0 ? (double)A->ctx->opt.wave_l2_pct : 80.0) / 100.0;
    if (window > budget * (double)A->ctx->prop.l2CacheSize) {
        *why = "wavefront window exceeds the L2 budget";
        WavePlan rej;  // remember the refusal: planning costs O(tiles) host work
        rej.k = k; rej.t_nnz = V.t_nnz; rej.t_rows = V.t_rows; rej.level_rows = lr; rej.slack = slack; rej.rejected = true;
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

    WavePlan p;
    p.k = k; p.t_nnz = V.t_nnz; p.t_rows = V.t_rows; p.level_rows = lr; p.slack = slack;
    p.ntasks = (int)tasks.size(); p.ngroups = ngroups; p.D = D;
    if (cudaMalloc(&p.d_tasks, sizeof(WaveTask) * tasks.size()) != cudaSuccess ||
        cudaMalloc(&p.d_counters, sizeof(int) * ((size_t)k * ngroups + 4)) != cudaSuccess ||
        cudaMalloc(&p.d_group_size, sizeof(int) * (size_t)k * ngroups) != cudaSuccess) {
        *why = "plan allocation failed";
        return nullptr;
    }
    cudaMemcpy(p.d_tasks, tasks.data(), sizeof(WaveTask) * tasks.size(), cudaMemcpyHostToDevice);
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
    *slack = (int)((pct / 100.0) * (double)(*grid_max) * (V.stages + 1) / (double)k + 0.999);  // +1: the pre-claimed item
    *fn_out = fn;
    *smem_out = smem;
    return NSK_OK;
}

bool nsk_mpk_wavefront_applicable(nsk_csr_t A, int k)
{
    if (k < 2 || A->mean_row > 12.0 || A->n == 0) return false;
    const char *why = nullptr;
    wave_fn fn; int smem, grid_max, slack;
    const int variant = wave_variant(A->ctx);
    if (wave_launch_shape(A->ctx, variant, false, k, &fn, &smem, &grid_max, &slack) != NSK_OK) return false;
    return get_plan(A, k, nullptr, g_wvariants[variant], slack, &why) != nullptr;
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
    WavePlan *plan = get_plan(A, k, level_rows, V, slack, &why);
    if (!plan) {
        nsk_set_error(ctx, "wavefront matrix powers not applicable: %s", why);
        return NSK_ERR_UNSUPPORTED;
    }
    const nsk_tiling *Tp = nullptr;
    NSK_TRY(nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp));
    const int grid = std::min(plan->ntasks, grid_max);

    const size_t ncnt = (size_t)k * plan->ngroups;
    NSK_CUDA(ctx, cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (ncnt + 4), ctx->stream));
    WaveParams P;
    P.tiles = Tp->d_tiles;
    P.tasks = plan->d_tasks;
    P.ntasks = plan->ntasks;
    P.ngroups = plan->ngroups;
    P.counters = plan->d_counters;
    P.group_size = plan->d_group_size;
    P.next = reinterpret_cast<unsigned int *>(plan->d_counters + ncnt);
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
