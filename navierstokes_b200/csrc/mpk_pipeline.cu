// mpk_pipeline.cu -- fused matrix-powers kernel, level-pipelined: levels[l] = A^(l+1) x, l = 0..k-1, in ONE
// persistent launch whose CTAs are specialised BY LEVEL.
//
// Replaces the reference's fused first-touch kernels SpM2V_CSR* (mpk/SpM2V.cpp:80-332), SpM3V and SpM4V
// (mpk/SpMVmulti0.cpp:132-221).  Same result as k row-sequential products (bit for bit in the exact modes:
// the per-row chain is consume_tile(), shared with every other kernel); only the schedule differs.
//
// Schedule.  The resident CTAs are split into k teams of G; team l computes power l+1 only, its CTA c taking
// tiles c, c+G, c+2G, ... of that level in global row order -- k concurrent streaming products, chained:
//   * forward dependency: tile t of level l starts when the level l-1 tile groups covering its column range
//     are complete (per-group completion counters in global memory, polled by a dependency warp that keeps a
//     watermark, so most items open without any memory traffic);
//   * back-pressure: level l may not run more than `lead` tiles ahead of level l+1, which bounds the
//     wavefront window ((k-1) * lead tiles of matrix data + level vectors) so it STAYS in the 126 MB L2:
//     only team 0 streams the operator from HBM, teams 1..k-1 re-read it from L2.
// Nothing is claimed dynamically and no level is ever blocked behind another level's item (the first fused
// kernel, mpk_wavefront.cu, interleaved levels in one in-order ring per CTA and needed a static skew
// large enough to hide every latency: its window overflowed L2 -- 3.8 GB of HBM traffic for 2.1 GB
// compulsory, profiles/r01_ncu_mpk_wavefront_c_summary.txt).
//
// Deadlock freedom: all CTAs are co-resident (grid <= occupancy); team k-1 has no back-pressure and team 0 no
// forward dependency; lead > reach, so the oldest unfinished tile of the last team that is behind always has
// its inputs complete.  Every spin is bounded and traps instead of hanging.
#include <algorithm>
#include <map>
#include <mutex>

#include "nsk_internal.h"
#include "ptx_helpers.cuh"
#include "stream_common.cuh"
#include "wave_common.h"

using namespace nskptx;

struct PipeItem {               // 32 bytes = two 16-byte loads
    int row0, nrows, nz0, nz1;  // the tile
    int pos;                    // position in global row order (completion group = pos / WF_GROUP)
    int ghi;                    // forward: groups [0, ghi] of level l-1 must be complete
    int gback;                  // back-pressure: groups [0, gback] of level l+1 must be complete (-1: none)
    int pad;
};

struct PipeParams {
    const PipeItem *items[NSK_MAX_K];  // per level, ascending position
    int count[NSK_MAX_K];
    int *counters;          // [k][ngroups], zeroed before the launch
    const int *group_size;  // [k][ngroups] tiles that report per group and level
    int ngroups;
    const int *ptrow;
    const int *indcol;
    const double *coef;
    const double *x;
    double *levels[NSK_MAX_K];
    int level_rows[NSK_MAX_K];
    int k;
    int team;       // CTAs per level
    int interleave; // 1: level = blockIdx % k, 0: level = blockIdx / team
};

// Advance the watermark w (all groups < w complete for `cnt`) until it passes `upto`; one warp, one round trip
// per 32 groups.  ld.acquire.gpu (LDG.STRONG.GPU + CCTL.IVALL) also drops this SM's stale L1 lines.
__device__ __forceinline__ int pipe_wait_groups(const int *cnt, const int *need, int ngroups, int w, int upto, int lane)
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
            __nanosleep(40);
            if (++spins > (1u << 24)) __trap();  // protocol bug: fail the launch, never hang
        }
    }
    return w;
}

template <int T_NNZ, int T_ROWS, int STAGES, int NCW, int MINB, int RPT, bool MULADD>
__global__ void __launch_bounds__((NCW + 2) * 32, MINB) mpk_pipeline_kernel(const PipeParams P)
{
    using Geo = StageGeom<T_NNZ, T_ROWS>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage_base = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Geo::BYTES * STAGES);  // TMA bytes of the stage landed
    uint64_t *ready = full + STAGES;                                            // inputs of the item are complete
    uint64_t *done = ready + STAGES;                                            // all consumer warps stored their rows

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], 1);
            mbar_init(&done[s], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int level = P.interleave ? (int)(blockIdx.x % P.k) : (int)(blockIdx.x / P.team);
    const int c = P.interleave ? (int)(blockIdx.x / P.k) : (int)(blockIdx.x % P.team);
    const int G = P.team;
    const int count = P.count[level];
    const int n_my = c < count ? (count - c + G - 1) / G : 0;
    const int4 *my = reinterpret_cast<const int4 *>(P.items[level]);  // item i of this CTA = my[2 * (c + i * G)]

    if (warp == NCW) {
        // ===== producer + publisher (one lane).  A stage is refilled the moment its previous item is done
        // (the slice depends on no flag); finished items are then published with ONE gpu-scope fence for all
        // items found finished at that moment, off the consumers' path. =====
        if (lane != 0) return;
        int pubgrp[STAGES];
        int it_load = 0, it_pub = 0;
        int4 da = make_int4(0, 0, 0, 0), db = da, na = da, nb = da;
        if (n_my > 0) { da = my[2 * (size_t)c]; db = my[2 * (size_t)c + 1]; }
        if (n_my > 1) { na = my[2 * (size_t)(c + G)]; nb = my[2 * (size_t)(c + G) + 1]; }
        auto load_next = [&]() {
            const int s = it_load % STAGES;
            unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
            int *hdr = reinterpret_cast<int *>(st + Geo::HDR_OFF);
            hdr[0] = da.x; hdr[1] = da.y; hdr[2] = da.z; hdr[3] = da.w;
#pragma unroll
            for (int j = 0; j < STAGES; j++)
                if (j == s) pubgrp[j] = db.x / WF_GROUP;
            const int row0 = da.x, nrows = da.y, nz0 = da.z, nz1 = da.w;
            const int a0 = nz0 & ~3, v0 = nz0 & ~1, p0 = row0 & ~3;
            const uint32_t cbytes = (uint32_t)(((nz1 - a0) + 3) & ~3) * 4u;
            const uint32_t vbytes = (uint32_t)(((nz1 - v0) + 1) & ~1) * 8u;
            const uint32_t pbytes = (uint32_t)(((row0 + nrows + 1 - p0) + 3) & ~3) * 4u;
            mbar_arrive_expect_tx(&full[s], cbytes + vbytes + pbytes);
            bulk_g2s(st + Geo::PTR_OFF, P.ptrow + p0, pbytes, &full[s]);
            if (cbytes) bulk_g2s(st + Geo::COL_OFF, P.indcol + a0, cbytes, &full[s]);
            if (vbytes) bulk_g2s(st + Geo::VAL_OFF, P.coef + v0, vbytes, &full[s]);
            ++it_load;
            da = na; db = nb;
            if (it_load + 1 < n_my) {
                const size_t q = (size_t)c + (size_t)(it_load + 1) * G;
                na = my[2 * q]; nb = my[2 * q + 1];
            }
        };
        while (it_load < n_my && it_load < STAGES) load_next();
        int *cnt = P.counters + (size_t)level * P.ngroups;
        while (it_pub < n_my) {
            int grp[STAGES];
            int n = 0;
#pragma unroll
            for (int j = 0; j < STAGES; j++) {
                if (j > n || it_pub + j >= n_my) continue;  // stop at the first unfinished item
                const int s = (it_pub + j) % STAGES;
                const uint32_t ph = ((it_pub + j) / STAGES) & 1;
                if (j == 0) mbar_wait(&done[s], ph);
                else if (!mbar_try_wait(&done[s], ph)) continue;
#pragma unroll
                for (int u = 0; u < STAGES; u++)
                    if (u == s) grp[j] = pubgrp[u];
                if (it_load < n_my) load_next();  // it_load == it_pub + j + STAGES: same stage
                n = j + 1;
            }
            __threadfence();
#pragma unroll
            for (int j = 0; j < STAGES; j++)
                if (j < n) red_relaxed_gpu_add(cnt + grp[j], 1);
            it_pub += n;
        }
        return;
    }

    if (warp == NCW + 1) {
        // ===== dependency warp: opens item `it` for the consumers once its forward inputs (level - 1) are
        // complete and the next level (level + 1) is not more than `lead` tiles behind. =====
        const bool fwd = level > 0, back = level < P.k - 1;
        const int *cnt_f = P.counters + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *need_f = P.group_size + (size_t)(fwd ? level - 1 : 0) * P.ngroups;
        const int *cnt_b = P.counters + (size_t)(back ? level + 1 : 0) * P.ngroups;
        const int *need_b = P.group_size + (size_t)(back ? level + 1 : 0) * P.ngroups;
        int wf = 0, wb = 0;
        int2 mine = make_int2(0, -1);
        for (int it = 0; it < n_my; ++it) {
            const int j = it & 31;
            if (j == 0) {  // lane u fetches the limits of item it + u
                mine = make_int2(0, -1);
                if (it + lane < n_my) {
                    const int4 b = my[2 * ((size_t)c + (size_t)(it + lane) * G) + 1];
                    mine = make_int2(b.y, b.z);
                }
            }
            const int ghi = __shfl_sync(0xffffffffu, mine.x, j);
            const int gback = __shfl_sync(0xffffffffu, mine.y, j);
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&done[s], ((it / STAGES) - 1) & 1);  // every consumer is past ready[s] of item it - STAGES
            if (back && gback >= wb) wb = pipe_wait_groups(cnt_b, need_b, P.ngroups, wb, gback, lane);
            if (fwd && ghi >= wf) wf = pipe_wait_groups(cnt_f, need_f, P.ngroups, wf, ghi, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);  // release.cta: orders the acquires above before the consumers
        }
        return;
    }

    // ===== consumer warps: independent of one another (no CTA-wide barrier in the loop) =====
    constexpr int NCT = NCW * 32;
    const double *src = level == 0 ? P.x : P.levels[level - 1];
    double *dst = P.levels[level];
    const int row_end = P.level_rows[level];
    for (int it = 0; it < n_my; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&ready[s], ph);
        mbar_wait(&full[s], ph);
        unsigned char *st = stage_base + (size_t)s * Geo::BYTES;
        const int *hdr = reinterpret_cast<const int *>(st + Geo::HDR_OFF);
        const int row0 = hdr[0], nrows = hdr[1], nz0 = hdr[2];
        const double *val_s = reinterpret_cast<const double *>(st + Geo::VAL_OFF);
        const int *col_s = reinterpret_cast<const int *>(st + Geo::COL_OFF);
        const int *ptr_s = reinterpret_cast<const int *>(st + Geo::PTR_OFF);
        const int vo = nz0 & ~1, co = nz0 & ~3, po = row0 & ~3;
        // level 0 reads x (constant for the launch: read-only path); level l >= 1 reads what other CTAs wrote
        // earlier in THIS launch: ordinary coherent loads, ordered by ready[s] after the dependency warp's
        // acquires (which also invalidated this SM's L1)
        if (level == 0)
            consume_tile<MULADD, true, RPT, NCT>(val_s, col_s, ptr_s, vo, co, po, row0, nrows, row_end, src, dst, tid);
        else
            consume_tile<MULADD, false, RPT, NCT>(val_s, col_s, ptr_s, vo, co, po, row0, nrows, row_end, src, dst, tid);
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[s]);  // release.cta; no GPU-scope fence on this path
    }
}

// -----------------------------------------------------------------------------------------------
// host side: plan (cached per operator / k / level_rows / geometry) and launch
// -----------------------------------------------------------------------------------------------
struct PipePlan {
    int k = 0, t_nnz = 0, t_rows = 0, team = 0, lead_pct = 0, l2_pct = 0;
    bool rejected = false;
    std::vector<int> level_rows;
    int ngroups = 0, reach = 0, lead = 0;
    std::vector<int> count;
    std::vector<size_t> item_off;
    PipeItem *d_items = nullptr;  // all levels back to back
    int *d_counters = nullptr;
    int *d_group_size = nullptr;
};

static std::map<nsk_csr_t, std::vector<PipePlan>> g_pipe;  // map guarded; an entry belongs to its operator's thread
static std::mutex g_pipe_mu;

void nsk_pipe_free(nsk_csr_t A)
{
    std::vector<PipePlan> mine;
    {
        std::lock_guard<std::mutex> lk(g_pipe_mu);
        auto it = g_pipe.find(A);
        if (it == g_pipe.end()) return;
        mine.swap(it->second);
        g_pipe.erase(it);
    }
    for (PipePlan &p : mine) {
        if (p.d_items) cudaFree(p.d_items);
        if (p.d_counters) cudaFree(p.d_counters);
        if (p.d_group_size) cudaFree(p.d_group_size);
    }
}

struct PipeVariant {
    int t_nnz, t_rows, stages, ncw, minb, rpt;
};
//                 T_NNZ T_ROWS STAGES NCW MINB RPT(rows per consumer thread and pass)
#define NSK_PIPE_VARIANTS(X) \
    X(0, 1792, 256, 2, 8, 4, 1)   \
    X(1, 1792, 256, 3, 8, 3, 1)   \
    X(2, 3584, 512, 2, 8, 2, 2)   \
    X(3, 3584, 512, 2, 16, 2, 1)  \
    X(4, 1344, 192, 3, 6, 4, 1)   \
    X(5, 3584, 512, 4, 16, 1, 1)  \
    X(6, 1792, 256, 4, 8, 2, 1)   \
    X(7, 896, 128, 4, 4, 4, 1)    \
    X(8, 3584, 512, 1, 16, 4, 1)  \
    X(9, 1792, 256, 2, 4, 4, 2)

static const PipeVariant g_pvariants[] = {
#define X(id, t, r, s, w, b, u) {t, r, s, w, b, u},
    NSK_PIPE_VARIANTS(X)
#undef X
};
static const int g_npvariants = sizeof(g_pvariants) / sizeof(g_pvariants[0]);

typedef void (*pipe_fn)(const PipeParams);
static pipe_fn pipe_lookup(int variant, bool muladd, int *smem)
{
    switch (variant) {
#define X(id, t, r, s, w, b, u)                               \
    case id:                                                  \
        *smem = StageGeom<t, r>::BYTES * s + 3 * s * 8 + 128; \
        return muladd ? mpk_pipeline_kernel<t, r, s, w, b, u, true> : mpk_pipeline_kernel<t, r, s, w, b, u, false>;
        NSK_PIPE_VARIANTS(X)
#undef X
    }
    return nullptr;
}

static int pipe_variant(nsk_ctx_t ctx)
{
    int v = (int)ctx->opt.pipe_variant - 1;  // option value 0 = default; n >= 1 selects table entry n - 1
    if (v < 0 || v >= g_npvariants) v = 0;
    return v;
}

static int pipe_launch_shape(nsk_ctx_t ctx, int variant, bool muladd, int k, pipe_fn *fn_out, int *smem_out, int *team)
{
    const PipeVariant &V = g_pvariants[variant];
    int smem = 0;
    pipe_fn fn = pipe_lookup(variant, muladd, &smem);
    NSK_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    NSK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, (V.ncw + 2) * 32, smem));
    if (ctx->opt.spmv_ctas_per_sm > 0) per_sm = std::min(per_sm, (int)ctx->opt.spmv_ctas_per_sm);
    const int resident = ctx->prop.multiProcessorCount * per_sm;
    *team = resident / k;
    *fn_out = fn;
    *smem_out = smem;
    return NSK_OK;
}

// lead: how many tiles level l may run ahead of level l+1 = reach + 1 + slack; the slack (default: two rounds
// of the team, i.e. what is in flight in its stage rings) absorbs the publish -> poll latency (~2 us).
static PipePlan *pipe_plan(nsk_csr_t A, int k, const int *level_rows, const PipeVariant &V, int team, const char **why)
{
    nsk_ctx_t ctx = A->ctx;
    g_pipe_mu.lock();
    std::vector<PipePlan> &plans = g_pipe[A];
    g_pipe_mu.unlock();
    std::vector<int> lr(k);
    for (int l = 0; l < k; l++) lr[l] = level_rows ? level_rows[l] : A->n;
    const int lead_pct = ctx->opt.wave_slack_pct >= 0 ? (int)ctx->opt.wave_slack_pct : 100;
    for (PipePlan &p : plans)
        if (p.k == k && p.t_nnz == V.t_nnz && p.t_rows == V.t_rows && p.team == team && p.level_rows == lr &&
            p.lead_pct == lead_pct && p.l2_pct == (int)ctx->opt.wave_l2_pct) {
            if (p.rejected) { *why = "wavefront window exceeds the L2 budget"; return nullptr; }
            return &p;
        }

    const nsk_tiling *Tp = nullptr;
    if (nsk_get_tiling(A, V.t_nnz, V.t_rows, &Tp) != NSK_OK) { *why = "tiling failed"; return nullptr; }
    const nsk_tiling &T = *Tp;
    WaveDeps D;
    if (!nsk_wave_deps(A, T, D, why)) return nullptr;
    const int ntiles = D.ntiles, ngroups = D.ngroups;
    const int slack = (int)((double)lead_pct / 100.0 * 2.0 * team + 0.999);
    const int lead = D.reach + 1 + slack + WF_GROUP;  // + one group: completion is only visible per group

    PipePlan p;
    p.k = k; p.t_nnz = V.t_nnz; p.t_rows = V.t_rows; p.team = team; p.lead_pct = lead_pct; p.level_rows = lr;
    p.l2_pct = (int)ctx->opt.wave_l2_pct;
    p.ngroups = ngroups; p.reach = D.reach; p.lead = lead;
    // The window that must stay in L2: (k-1)*lead tiles of matrix data plus the level vectors over it.
    const double tile_bytes = (12.0 * (double)A->nnz + 8.0 * (double)A->n * (k + 1)) / (double)ntiles;  // mean tile
    const double window = (double)(k - 1) * lead * tile_bytes;
    const double budget = (ctx->opt.wave_l2_pct > 0 ? (double)ctx->opt.wave_l2_pct : 80.0) / 100.0;
    if (window > budget * (double)ctx->prop.l2CacheSize) {
        *why = "wavefront window exceeds the L2 budget";
        p.rejected = true;  // remember the refusal: planning costs O(tiles) host work
        plans.push_back(p);
        return nullptr;
    }

    std::vector<PipeItem> items;
    items.reserve((size_t)k * ntiles);
    std::vector<int> gsize((size_t)k * ngroups, 0);
    p.count.assign(k, 0);
    p.item_off.assign(k, 0);
    for (int l = 0; l < k; l++) {
        p.item_off[l] = items.size();
        for (int pos = 0; pos < ntiles; pos++) {
            const int t = D.tile_at_pos[pos];
            const nsk_tile &tl = T.h_tiles[t];
            if (tl.row0 >= lr[l]) continue;  // outside this level's row prefix (distributed shrink)
            int gback = -1;
            if (l < k - 1 && pos - lead >= 0) gback = (pos - lead) / WF_GROUP - 1;  // groups wholly below pos - lead
            items.push_back(PipeItem{tl.row0, tl.nrows, tl.nz0, tl.nz1, pos, D.ghi[t], gback, 0});
            gsize[(size_t)l * ngroups + pos / WF_GROUP]++;
            p.count[l]++;
        }
    }
    if (cudaMalloc(&p.d_items, sizeof(PipeItem) * (items.size() + 1)) != cudaSuccess ||
        cudaMalloc(&p.d_counters, sizeof(int) * ((size_t)k * ngroups + 4)) != cudaSuccess ||
        cudaMalloc(&p.d_group_size, sizeof(int) * (size_t)k * ngroups) != cudaSuccess) {
        *why = "plan allocation failed";
        return nullptr;
    }
    cudaMemcpy(p.d_items, items.data(), sizeof(PipeItem) * items.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_group_size, gsize.data(), sizeof(int) * gsize.size(), cudaMemcpyHostToDevice);
    plans.push_back(p);
    return &plans.back();
}

bool nsk_mpk_pipeline_applicable(nsk_csr_t A, int k)
{
    if (k < 2 || A->mean_row > 64.0 || A->n == 0) return false;
    const char *why = nullptr;
    pipe_fn fn; int smem, team;
    const int variant = pipe_variant(A->ctx);
    if (pipe_launch_shape(A->ctx, variant, false, k, &fn, &smem, &team) != NSK_OK || team < 1) return false;
    return pipe_plan(A, k, nullptr, g_pvariants[variant], team, &why) != nullptr;
}

int nsk_mpk_pipeline(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                     const int *level_rows)
{
    nsk_ctx_t ctx = A->ctx;
    const int variant = pipe_variant(ctx);
    const PipeVariant &V = g_pvariants[variant];
    const char *why = "";
    pipe_fn fn; int smem = 0, team = 0;
    NSK_TRY(pipe_launch_shape(ctx, variant, mode == NSK_EXACT_MULADD, k, &fn, &smem, &team));
    if (team < 1) {
        nsk_set_error(ctx, "level-pipelined matrix powers not applicable: fewer resident CTAs than levels");
        return NSK_ERR_UNSUPPORTED;
    }
    PipePlan *plan = pipe_plan(A, k, level_rows, V, team, &why);
    if (!plan) {
        nsk_set_error(ctx, "level-pipelined matrix powers not applicable: %s", why);
        return NSK_ERR_UNSUPPORTED;
    }
    const size_t ncnt = (size_t)k * plan->ngroups;
    NSK_CUDA(ctx, cudaMemsetAsync(plan->d_counters, 0, sizeof(int) * (ncnt + 4), ctx->stream));
    PipeParams P;
    for (int l = 0; l < NSK_MAX_K; l++) {
        P.items[l] = l < k ? plan->d_items + plan->item_off[l] : nullptr;
        P.count[l] = l < k ? plan->count[l] : 0;
        P.levels[l] = l < k ? d_levels[l] : nullptr;
        P.level_rows[l] = l < k ? plan->level_rows[l] : 0;
    }
    P.counters = plan->d_counters;
    P.group_size = plan->d_group_size;
    P.ngroups = plan->ngroups;
    P.ptrow = A->d_ptrow;
    P.indcol = A->d_indcol;
    P.coef = A->d_coef;
    P.x = d_x;
    P.k = k;
    P.team = team;
    P.interleave = ctx->opt.pipe_interleave ? 1 : 0;
    fn<<<team * k, (V.ncw + 2) * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}
