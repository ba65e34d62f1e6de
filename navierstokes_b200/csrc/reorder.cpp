// reorder.cpp -- bandwidth-reducing ordering of a CSR operator on the host: reverse Cuthill-McKee + symmetric
// permutation.  The reference has no partitioner or ordering code at all; the north star asks for
// "METIS-style or RCM-ordered blocks": a banded operator is what makes contiguous row slabs a good partition (halo =
// the band) and what keeps the matrix-powers window (reach of a tile) small.
//
// nsk_rcm           perm[new] = old.  Level-set BFS from a pseudo-peripheral node of every connected component
//                   (George-Liu), neighbours visited in order of ascending degree, the whole order reversed.
//                   Works on the pattern of A + A^T (structural symmetry is not assumed).
// nsk_csr_permute   B = P A P^T with every row's columns ascending (the values travel with their entries).
// nsk_csr_bandwidth max |i - j| over the entries.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <thread>
#include <vector>

#include "../../include/nsk.h"

#define NSK_API extern "C" __attribute__((visibility("default")))

namespace {

// rows [0, n) split over the host threads
template <class F>
void parallel_rows(int n, F body)
{
    const int nth = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (n < 100000 || nth == 1) { body(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nth; t++) {
        const int a = (int)((int64_t)n * t / nth), b = (int)((int64_t)n * (t + 1) / nth);
        th.emplace_back(body, a, b);
    }
    for (auto &x : th) x.join();
}

// adjacency of A + A^T without the diagonal, neighbours sorted and unique
void symmetric_pattern(int n, const int *ptrow, const int *indcol, std::vector<int64_t> &ptr, std::vector<int> &adj)
{
    std::vector<int64_t> cnt((size_t)n + 1, 0);
    for (int i = 0; i < n; i++)
        for (int j = ptrow[i]; j < ptrow[i + 1]; j++) {
            const int c = indcol[j];
            if (c == i || c < 0 || c >= n) continue;
            cnt[(size_t)i + 1]++;
            cnt[(size_t)c + 1]++;
        }
    ptr.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) ptr[(size_t)i + 1] = ptr[i] + cnt[(size_t)i + 1];
    std::vector<int> raw((size_t)ptr[n]);
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < n; i++)
        for (int j = ptrow[i]; j < ptrow[i + 1]; j++) {
            const int c = indcol[j];
            if (c == i || c < 0 || c >= n) continue;
            raw[(size_t)fill[i]++] = c;
            raw[(size_t)fill[c]++] = i;
        }
    std::vector<int64_t> nptr((size_t)n + 1, 0);
    parallel_rows(n, [&](int a, int b) {  // sort every row's list, unique in place, remember the count
        for (int i = a; i < b; i++) {
            int *first = raw.data() + ptr[i], *last = raw.data() + ptr[(size_t)i + 1];
            std::sort(first, last);
            nptr[(size_t)i + 1] = std::unique(first, last) - first;
        }
    });
    for (int i = 0; i < n; i++) nptr[(size_t)i + 1] += nptr[i];
    adj.resize((size_t)nptr[n]);
    parallel_rows(n, [&](int a, int b) {
        for (int i = a; i < b; i++)
            std::copy(raw.data() + ptr[i], raw.data() + ptr[i] + (nptr[(size_t)i + 1] - nptr[i]), adj.data() + nptr[i]);
    });
    ptr.swap(nptr);
}

// BFS from `root` over unvisited-by-`mark` nodes of its component; fills `order` (level by level) and returns the
// number of levels; level_start gets the offsets of the levels in `order`.
int bfs_levels(int root, const std::vector<int64_t> &ptr, const std::vector<int> &adj, std::vector<int> &stamp, int tag,
               std::vector<int> &order, std::vector<int> &level_start)
{
    order.clear();
    level_start.clear();
    order.push_back(root);
    stamp[root] = tag;
    size_t head = 0;
    int levels = 0;
    while (head < order.size()) {
        level_start.push_back((int)head);
        const size_t end = order.size();
        for (; head < end; head++) {
            const int u = order[head];
            for (int64_t p = ptr[u]; p < ptr[(size_t)u + 1]; p++) {
                const int v = adj[(size_t)p];
                if (stamp[v] != tag) {
                    stamp[v] = tag;
                    order.push_back(v);
                }
            }
        }
        levels++;
    }
    level_start.push_back((int)order.size());
    return levels;
}

}  // namespace

NSK_API int nsk_rcm(int n, const int *ptrow, const int *indcol, int *perm)
{
    if (n < 0 || !ptrow || !perm || (n > 0 && ptrow[n] > 0 && !indcol)) return NSK_ERR_INVALID;
    if (n == 0) return NSK_OK;
    std::vector<int64_t> ptr;
    std::vector<int> adj;
    symmetric_pattern(n, ptrow, indcol, ptr, adj);
    std::vector<int> deg(n);
    for (int i = 0; i < n; i++) deg[i] = (int)(ptr[(size_t)i + 1] - ptr[i]);
    std::vector<int> stamp(n, 0), order, level_start, nbr;
    std::vector<char> done(n, 0);
    std::vector<int> cm;  // Cuthill-McKee order, all components
    cm.reserve(n);
    // components are started from their lowest-degree unvisited node, scanned in index order
    std::vector<int> by_degree(n);
    std::iota(by_degree.begin(), by_degree.end(), 0);
    std::stable_sort(by_degree.begin(), by_degree.end(), [&](int a, int b) { return deg[a] < deg[b]; });
    int tag = 0;
    for (int s : by_degree) {
        if (done[s]) continue;
        // pseudo-peripheral node: repeat BFS from a lowest-degree node of the last level while the depth grows
        int root = s, depth = -1;
        for (;;) {
            const int levels = bfs_levels(root, ptr, adj, stamp, ++tag, order, level_start);
            if (levels <= depth) break;
            depth = levels;
            int best = -1;
            for (int p = level_start[levels - 1]; p < level_start[levels]; p++)
                if (best < 0 || deg[order[p]] < deg[best]) best = order[p];
            if (best == root) break;
            root = best;
        }
        // Cuthill-McKee from root: queue order, neighbours appended by ascending degree (ties by index: adj is sorted)
        const size_t first = cm.size();
        cm.push_back(root);
        done[root] = 1;
        for (size_t head = first; head < cm.size(); head++) {
            const int u = cm[head];
            nbr.clear();
            for (int64_t p = ptr[u]; p < ptr[(size_t)u + 1]; p++) {
                const int v = adj[(size_t)p];
                if (!done[v]) {
                    done[v] = 1;
                    nbr.push_back(v);
                }
            }
            std::stable_sort(nbr.begin(), nbr.end(), [&](int a, int b) { return deg[a] < deg[b]; });
            cm.insert(cm.end(), nbr.begin(), nbr.end());
        }
    }
    for (int i = 0; i < n; i++) perm[i] = cm[(size_t)(n - 1 - i)];
    return NSK_OK;
}

NSK_API int nsk_csr_permute(int n, const int *ptrow, const int *indcol, const double *coef, const int *perm, int *ptrow_out,
                            int *indcol_out, double *coef_out)
{
    if (n < 0 || !ptrow || !perm || !ptrow_out) return NSK_ERR_INVALID;
    std::vector<int> inv(n, -1);
    for (int i = 0; i < n; i++) {
        if (perm[i] < 0 || perm[i] >= n || inv[perm[i]] >= 0) return NSK_ERR_INVALID;  // not a permutation
        inv[perm[i]] = i;
    }
    ptrow_out[0] = 0;
    for (int i = 0; i < n; i++) ptrow_out[i + 1] = ptrow_out[i] + (ptrow[perm[i] + 1] - ptrow[perm[i]]);
    for (int i = 0; i < n; i++)
        for (int j = ptrow[i]; j < ptrow[i + 1]; j++)
            if (indcol[j] < 0 || indcol[j] >= n) return NSK_ERR_INVALID;
    parallel_rows(n, [&](int a, int b) {
        std::vector<std::pair<int, double>> row;
        for (int i = a; i < b; i++) {
            const int o = perm[i];
            row.clear();
            for (int j = ptrow[o]; j < ptrow[o + 1]; j++) row.emplace_back(inv[indcol[j]], coef ? coef[j] : 0.0);
            std::stable_sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
                return x.first < y.first;
            });
            int q = ptrow_out[i];
            for (const auto &e : row) {
                indcol_out[q] = e.first;
                if (coef_out) coef_out[q] = e.second;
                q++;
            }
        }
    });
    return NSK_OK;
}

NSK_API int64_t nsk_csr_bandwidth(int n, const int *ptrow, const int *indcol)
{
    int64_t bw = 0;
    for (int i = 0; i < n; i++)
        for (int j = ptrow[i]; j < ptrow[i + 1]; j++) bw = std::max<int64_t>(bw, indcol[j] > i ? indcol[j] - i : i - indcol[j]);
    return bw;
}
