// dist_plan.cpp -- host-only planning of the row-partitioned operator: ghost rings, local
// renumbering, per-peer exchange lists (nsk_plan_* of include/nsk.h).  No CUDA, no communicator:
// it runs on a CPU-only box and is what the world_size-2 gloo tests exercise.
//
// The reference is single-process; this introduces the partitioning the north star
// asks for.  Scheme (PA1 of the matrix-powers literature): a rank that owns rows R and wants k
// powers needs x on N_k(R), the k-step neighbourhood of R in the graph of A, and recomputes level l
// on N_{k-l}(R).  Rings N_j \ N_{j-1} are stored after the owned rows, ascending by global id, so a
// level is always a PREFIX of the local rows and a peer's contribution to a ring is one contiguous
// range of the local vector.
#include <algorithm>
#include <map>
#include <string.h>

#include "nsk_internal.h"
#include "dist_plan.h"

NSK_API int nsk_plan_create(int nranks, int rank, const int *row_starts, int depth, nsk_plan_t *out)
{
    if (!out || !row_starts || nranks < 1 || rank < 0 || rank >= nranks || depth < 1 || depth > NSK_MAX_K) {
        nsk_set_error(nullptr, "nsk_plan_create: bad arguments");
        return NSK_ERR_INVALID;
    }
    for (int r = 0; r < nranks; r++)
        if (row_starts[r + 1] < row_starts[r]) {
            nsk_set_error(nullptr, "nsk_plan_create: row_starts must be non-decreasing");
            return NSK_ERR_INVALID;
        }
    nsk_plan_s *p = new nsk_plan_s();
    p->nranks = nranks;
    p->rank = rank;
    p->depth = depth;
    p->row_starts.assign(row_starts, row_starts + nranks + 1);
    p->own_begin = row_starts[rank];
    p->own_end = row_starts[rank + 1];
    p->n_global = row_starts[nranks];
    p->rings.resize(depth + 1);
    p->frontier.resize(p->own_end - p->own_begin);
    for (int i = 0; i < (int)p->frontier.size(); i++) p->frontier[i] = p->own_begin + i;
    p->ptr.push_back(0);
    *out = p;
    return NSK_OK;
}

NSK_API int nsk_plan_destroy(nsk_plan_t p)
{
    delete p;
    return NSK_OK;
}

NSK_API int nsk_plan_frontier(nsk_plan_t p, int *count, const int **rows)
{
    if (!p || !count) return NSK_ERR_INVALID;
    if (p->stage >= p->depth) {
        *count = 0;
        if (rows) *rows = nullptr;
        return NSK_OK;
    }
    *count = (int)p->frontier.size();
    if (rows) *rows = p->frontier.data();
    return NSK_OK;
}

NSK_API int nsk_plan_add_rows(nsk_plan_t p, int count, const int *ptr, const int *cols, const double *vals)
{
    if (!p || p->finalized || p->stage >= p->depth || count != (int)p->frontier.size() || (count && !ptr)) {
        nsk_set_error(nullptr, "nsk_plan_add_rows: wrong stage or row count");
        return NSK_ERR_INVALID;
    }
    const int64_t add = count ? ptr[count] - ptr[0] : 0;
    if ((int64_t)p->cols.size() + add >= (int64_t)2147483647) {
        nsk_set_error(nullptr, "nsk_plan_add_rows: local nnz exceeds int32");
        return NSK_ERR_INVALID;
    }
    const int base = (int)p->cols.size();
    for (int i = 0; i < count; i++) p->ptr.push_back(base + (ptr[i + 1] - ptr[0]));
    std::vector<int> cand;
    for (int64_t e = ptr ? ptr[0] : 0; e < (ptr ? ptr[count] : 0); e++) {
        const int c = cols[e];
        if (c < 0 || c >= p->n_global) {
            nsk_set_error(nullptr, "nsk_plan_add_rows: column %d outside [0,%d)", c, p->n_global);
            return NSK_ERR_INVALID;
        }
        p->cols.push_back(c);
        p->vals.push_back(vals[e]);
        if (c < p->own_begin || c >= p->own_end) cand.push_back(c);
    }
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    // new ring = candidates not already known as ghosts
    std::vector<int> ring;
    std::set_difference(cand.begin(), cand.end(), p->known.begin(), p->known.end(), std::back_inserter(ring));
    p->stage++;
    p->rings[p->stage] = ring;
    std::vector<int> merged;
    merged.reserve(p->known.size() + ring.size());
    std::merge(p->known.begin(), p->known.end(), ring.begin(), ring.end(), std::back_inserter(merged));
    p->known.swap(merged);
    p->frontier = (p->stage < p->depth) ? ring : std::vector<int>();
    return NSK_OK;
}

NSK_API int nsk_plan_finalize(nsk_plan_t p)
{
    if (!p || p->stage != p->depth) {
        nsk_set_error(nullptr, "nsk_plan_finalize: rows of every ring have not been supplied yet");
        return NSK_ERR_INVALID;
    }
    if (p->finalized) return NSK_OK;
    const int n_owned = p->own_end - p->own_begin;
    p->ring_start.assign(p->depth + 2, 0);
    p->ring_start[1] = n_owned;
    for (int r = 1; r <= p->depth; r++) p->ring_start[r + 1] = p->ring_start[r] + (int)p->rings[r].size();
    // ghost lookup table: (global id -> local index), sorted by global id
    std::vector<std::pair<int, int>> lut;
    lut.reserve(p->known.size());
    p->ghost_gids.clear();
    for (int r = 1; r <= p->depth; r++)
        for (int i = 0; i < (int)p->rings[r].size(); i++) {
            lut.emplace_back(p->rings[r][i], p->ring_start[r] + i);
            p->ghost_gids.push_back(p->rings[r][i]);
        }
    std::sort(lut.begin(), lut.end());
    p->local_cols.resize(p->cols.size());
    for (size_t e = 0; e < p->cols.size(); e++) {
        const int c = p->cols[e];
        if (c >= p->own_begin && c < p->own_end) {
            p->local_cols[e] = c - p->own_begin;
        } else {
            auto it = std::lower_bound(lut.begin(), lut.end(), std::make_pair(c, -1));
            if (it == lut.end() || it->first != c) {
                nsk_set_error(nullptr, "nsk_plan_finalize: column %d has no local index", c);
                return NSK_ERR_INVALID;
            }
            p->local_cols[e] = it->second;
        }
    }
    // requests per owner rank: ghosts are ascending inside a ring and ownership is monotone in the id,
    // so each (ring, owner) pair is one contiguous slice
    p->req.clear();
    for (int r = 1; r <= p->depth; r++) {
        const std::vector<int> &ring = p->rings[r];
        size_t i = 0;
        while (i < ring.size()) {
            const int owner = (int)(std::upper_bound(p->row_starts.begin(), p->row_starts.end(), ring[i]) -
                                    p->row_starts.begin()) - 1;
            size_t j = i;
            while (j < ring.size() && ring[j] < p->row_starts[owner + 1]) j++;
            nsk_plan_s::Req &q = p->req[owner];
            if (q.ring_count.empty()) {
                q.ring_count.assign(p->depth, 0);
                q.ring_local_start.assign(p->depth, 0);
            }
            q.ring_count[r - 1] = (int)(j - i);
            q.ring_local_start[r - 1] = p->ring_start[r] + (int)i;
            q.gids.insert(q.gids.end(), ring.begin() + i, ring.begin() + j);
            i = j;
        }
    }
    p->finalized = true;
    return NSK_OK;
}

NSK_API int nsk_plan_sizes(nsk_plan_t p, int *n_owned, int *n_rows_local, int *n_cols_local, int64_t *nnz,
                           int *level_rows, int *ring_start)
{
    if (!p || !p->finalized) return NSK_ERR_INVALID;
    if (n_owned) *n_owned = p->own_end - p->own_begin;
    if (n_rows_local) *n_rows_local = p->ring_start[p->depth];
    if (n_cols_local) *n_cols_local = p->ring_start[p->depth + 1];
    if (nnz) *nnz = (int64_t)p->cols.size();
    if (level_rows)
        for (int l = 0; l < p->depth; l++) level_rows[l] = p->ring_start[p->depth - l];
    if (ring_start)
        for (int r = 0; r <= p->depth + 1; r++) ring_start[r] = p->ring_start[r];
    return NSK_OK;
}

NSK_API int nsk_plan_ghosts(nsk_plan_t p, const int **gids)
{
    if (!p || !p->finalized || !gids) return NSK_ERR_INVALID;
    *gids = p->ghost_gids.data();
    return NSK_OK;
}

NSK_API int nsk_plan_local_csr(nsk_plan_t p, const int **ptrow, const int **indcol, const double **coef)
{
    if (!p || !p->finalized) return NSK_ERR_INVALID;
    if (ptrow) *ptrow = p->ptr.data();
    if (indcol) *indcol = p->local_cols.data();
    if (coef) *coef = p->vals.data();
    return NSK_OK;
}

NSK_API int nsk_plan_requests(nsk_plan_t p, int peer, int *count, const int **gids, int *ring_counts)
{
    if (!p || !p->finalized || !count) return NSK_ERR_INVALID;
    auto it = p->req.find(peer);
    if (it == p->req.end()) {
        *count = 0;
        if (gids) *gids = nullptr;
        if (ring_counts) memset(ring_counts, 0, sizeof(int) * p->depth);
        return NSK_OK;
    }
    *count = (int)it->second.gids.size();
    if (gids) *gids = it->second.gids.data();
    if (ring_counts) memcpy(ring_counts, it->second.ring_count.data(), sizeof(int) * p->depth);
    return NSK_OK;
}

NSK_API int nsk_plan_add_send(nsk_plan_t p, int peer, int count, const int *gids, const int *ring_counts)
{
    if (!p || !p->finalized || peer < 0 || peer >= p->nranks || count < 0 || (count && (!gids || !ring_counts))) {
        nsk_set_error(nullptr, "nsk_plan_add_send: bad arguments");
        return NSK_ERR_INVALID;
    }
    if (count == 0) return NSK_OK;
    nsk_plan_s::Send s;
    s.ring_count.assign(ring_counts, ring_counts + p->depth);
    int total = 0;
    for (int r = 0; r < p->depth; r++) total += ring_counts[r];
    if (total != count) {
        nsk_set_error(nullptr, "nsk_plan_add_send: ring counts do not add up");
        return NSK_ERR_INVALID;
    }
    s.local_idx.resize(count);
    for (int i = 0; i < count; i++) {
        if (gids[i] < p->own_begin || gids[i] >= p->own_end) {
            nsk_set_error(nullptr, "nsk_plan_add_send: rank %d asked rank %d for row %d which it does not own", peer,
                          p->rank, gids[i]);
            return NSK_ERR_INVALID;
        }
        s.local_idx[i] = gids[i] - p->own_begin;
    }
    p->sends[peer] = s;
    return NSK_OK;
}
