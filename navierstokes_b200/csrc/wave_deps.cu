// wave_deps.cu -- host-side dependency geometry of the fused matrix-powers kernels (packed.cu, sell.cu): per operator
// the column extent of every 32-row block, and from it, for any tiling, the tile order in global row order and the
// range of tiles each tile's columns refer to.  Host only; no kernels.
#include <algorithm>
#include <climits>
#include <map>
#include <mutex>

#include <thread>

#include "nsk_internal.h"
#include "wave_common.h"

struct WaveState {
    // column extents per row block (<= 32 rows, never straddling a break), in GLOBAL-ORDER rank space
    std::vector<int> blk_row0, blk_min, blk_max;
};

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
    nsk_packed_free(A);
    nsk_sell_free(A);
    S.blk_row0.clear(); S.blk_min.clear(); S.blk_max.clear();
    const int n = A->n;
    const bool ranked = !A->row_rank.empty();
    // blocks of 32 rows (cut at the breaks of a distributed slab), then their column extents -- the second pass threaded
    size_t bi = 0;
    int r = 0;
    while (r < n) {
        while (bi < A->breaks.size() && A->breaks[bi] <= r) bi++;
        const int seg_end = bi < A->breaks.size() ? std::min(n, A->breaks[bi]) : n;
        S.blk_row0.push_back(r);
        r = std::min(seg_end, r + 32);
    }
    const int nblk = (int)S.blk_row0.size();
    S.blk_min.assign((size_t)nblk, INT32_MAX);
    S.blk_max.assign((size_t)nblk, -1);
    const int nth = nblk > (1 << 15) ? (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
    auto work = [&](int t) {
        const int b0 = (int)((int64_t)nblk * t / nth), b1 = (int)((int64_t)nblk * (t + 1) / nth);
        for (int b = b0; b < b1; b++) {
            const int ra = S.blk_row0[(size_t)b], rb = b + 1 < nblk ? S.blk_row0[(size_t)b + 1] : n;
            int mn = INT32_MAX, mx = -1;
            for (int j = ptrow[ra]; j < ptrow[rb]; j++) {
                int c = indcol[j];
                if (c >= n) continue;
                if (ranked) c = A->row_rank[c];
                mn = c < mn ? c : mn;
                mx = c > mx ? c : mx;
            }
            S.blk_min[(size_t)b] = mn;
            S.blk_max[(size_t)b] = mx;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nth; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
}

void nsk_wave_free(nsk_csr_t A)
{
    std::lock_guard<std::mutex> lk(g_wave_mu);
    g_wave.erase(A);
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

