// ingest.cpp -- the step in front of the hot path: Matrix Market text -> COO -> CSR / 4x4 block CSR with the
// reference's exact semantics, done fast on the host (no GPU involved; part of libnsk.so so the drop-in keeps one
// library).
//
//   nsk_mtx_read      <- the reader every reference driver inlines (mpk/SpM2V.cpp:815-852, same text in
//                        2SpMV.cpp / SpMVmulti*.cpp): first line skipped unconditionally, then '%' lines, a size
//                        line "rows cols nnz" (cols ignored, a `symmetric` banner ignored), nnz entries "i j v"
//                        1-based, v parsed into a FLOAT and widened.
//   nsk_coo2csr       <- COO2CSR + generate_CSR (mpk/utils.cpp:97-127, 5-43): columns ascending inside a row, a
//                        repeated (i,j) is DROPPED (first wins).  The reference inserts into one std::list per row,
//                        O(row length) per entry; here: stable counting sort by row, then a stable sort by column
//                        per row, rows in parallel -- same result, O(nnz log row).
//   nsk_coo2bcsr4     <- generate_BCSR4 (mpk/utils.cpp:45-95): nrow/4 block rows, block columns in
//                        FIRST-APPEARANCE order, row-major 4x4 blocks with explicit zeros, a repeated (i,j)
//                        OVERWRITES (last wins).  The reference searches the block list linearly per entry.
// Parity: tests/test_ingest.py against the fixtures produced by the compiled reference (tests/golden/formats.npz)
// and against the oracle's restatement on random inputs with duplicates.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "nsk_internal.h"

namespace {

template <class F>
void parallel_rows(int n, F &&body)
{
    const int nth = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 4096 || nth == 1) {
        body(0, n);
        return;
    }
    std::atomic<int> next(0);
    const int chunk = std::max(256, n / (nth * 16));
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(chunk);
            if (b >= n) return;
            body(b, std::min(n, b + chunk));
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nth; i++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

}  // namespace

// Returns the number of entries kept (<= nnz) or a negative status.  ptrow[nrow+1]; indcol/coef sized nnz.
NSK_API int64_t nsk_coo2csr(int nrow, int64_t nnz, const int *irow, const int *jcol, const double *val, int *ptrow,
                            int *indcol, double *coef)
{
    if (nrow < 0 || nnz < 0 || !ptrow || (nnz > 0 && (!irow || !jcol || !val || !indcol || !coef))) return NSK_ERR_INVALID;
    for (int64_t e = 0; e < nnz; e++)
        if (irow[e] < 0 || irow[e] >= nrow || jcol[e] < 0) return NSK_ERR_INVALID;
    // stable counting sort by row
    std::vector<int64_t> start((size_t)nrow + 1, 0);
    for (int64_t e = 0; e < nnz; e++) start[(size_t)irow[e] + 1]++;
    for (int i = 0; i < nrow; i++) start[(size_t)i + 1] += start[i];
    std::vector<int64_t> perm((size_t)std::max<int64_t>(nnz, 1));
    {
        std::vector<int64_t> pos(start.begin(), start.end() - 1);
        for (int64_t e = 0; e < nnz; e++) perm[(size_t)pos[irow[e]]++] = e;
    }
    // per row: stable sort by column (ties keep input order), keep the first of every run of equal columns
    std::vector<int> kept((size_t)nrow, 0);
    parallel_rows(nrow, [&](int r0, int r1) {
        for (int i = r0; i < r1; i++) {
            int64_t *b = perm.data() + start[i], *e = perm.data() + start[(size_t)i + 1];
            std::stable_sort(b, e, [&](int64_t x, int64_t y) { return jcol[x] < jcol[y]; });
            int64_t *w = b;
            for (int64_t *q = b; q < e; q++)
                if (q == b || jcol[*q] != jcol[*(q - 1)]) *w++ = *q;
            kept[i] = (int)(w - b);
        }
    });
    int64_t k = 0;
    ptrow[0] = 0;
    for (int i = 0; i < nrow; i++) {
        k += kept[i];
        if (k > 2147483647LL) return NSK_ERR_INVALID;
        ptrow[i + 1] = (int)k;
    }
    parallel_rows(nrow, [&](int r0, int r1) {
        for (int i = r0; i < r1; i++) {
            const int64_t *b = perm.data() + start[i];
            for (int q = 0; q < kept[i]; q++) {
                indcol[(size_t)ptrow[i] + q] = jcol[b[q]];
                coef[(size_t)ptrow[i] + q] = val[b[q]];
            }
        }
    });
    return k;
}

// Two-call protocol: indcol == nullptr returns the block count; then ptrow[nrow/4+1], indcol[nblk], coef[16*nblk].
NSK_API int64_t nsk_coo2bcsr4(int nrow, int64_t nnz, const int *irow, const int *jcol, const double *val, int *ptrow,
                              int *indcol, double *coef)
{
    if (nrow < 0 || nnz < 0 || (nnz > 0 && (!irow || !jcol || !val))) return NSK_ERR_INVALID;
    const int nb = nrow / 4;  // the reference uses nrow/4 block rows (utils.cpp:49); n must be a multiple of 4
    for (int64_t e = 0; e < nnz; e++)
        if (irow[e] < 0 || irow[e] / 4 >= nb || jcol[e] < 0) return NSK_ERR_INVALID;
    // entries grouped by block row, input order kept
    std::vector<int64_t> start((size_t)nb + 1, 0);
    for (int64_t e = 0; e < nnz; e++) start[(size_t)(irow[e] / 4) + 1]++;
    for (int b = 0; b < nb; b++) start[(size_t)b + 1] += start[b];
    std::vector<int64_t> perm((size_t)std::max<int64_t>(nnz, 1));
    {
        std::vector<int64_t> pos(start.begin(), start.end() - 1);
        for (int64_t e = 0; e < nnz; e++) perm[(size_t)pos[irow[e] / 4]++] = e;
    }
    // per block row: rank of each block column in order of first appearance.  slot[e] = block index inside the row.
    std::vector<int> slot((size_t)std::max<int64_t>(nnz, 1));
    std::vector<int> nblk((size_t)nb, 0);
    parallel_rows(nb, [&](int b0, int b1) {
        std::vector<std::pair<int, int64_t>> key;  // (block column, position in the row's input order)
        std::vector<int> first_rank;
        for (int bi = b0; bi < b1; bi++) {
            const int64_t lo = start[bi], hi = start[(size_t)bi + 1];
            const int64_t m = hi - lo;
            key.resize((size_t)m);
            for (int64_t q = 0; q < m; q++) key[(size_t)q] = {jcol[perm[(size_t)(lo + q)]] / 4, q};
            std::sort(key.begin(), key.end());
            // distinct block columns with their first position, then ranked by that position
            std::vector<std::pair<int64_t, int>> firsts;  // (first position, block column)
            for (int64_t q = 0; q < m; q++)
                if (q == 0 || key[(size_t)q].first != key[(size_t)q - 1].first) firsts.push_back({key[(size_t)q].second, key[(size_t)q].first});
            std::sort(firsts.begin(), firsts.end());
            nblk[bi] = (int)firsts.size();
            // block column -> rank: walk `key` (sorted by column) with a lookup over firsts sorted by column
            std::vector<std::pair<int, int>> col_rank(firsts.size());
            for (size_t r = 0; r < firsts.size(); r++) col_rank[r] = {firsts[r].second, (int)r};
            std::sort(col_rank.begin(), col_rank.end());
            size_t c = 0;
            for (int64_t q = 0; q < m; q++) {
                while (col_rank[c].first != key[(size_t)q].first) c++;
                slot[(size_t)(lo + key[(size_t)q].second)] = col_rank[c].second;
            }
        }
    });
    int64_t total = 0;
    for (int b = 0; b < nb; b++) total += nblk[b];
    if (!indcol) return total;
    if (!ptrow || !coef) return NSK_ERR_INVALID;
    ptrow[0] = 0;
    for (int b = 0; b < nb; b++) ptrow[b + 1] = ptrow[b] + nblk[b];
    memset(coef, 0, sizeof(double) * 16 * (size_t)total);
    parallel_rows(nb, [&](int b0, int b1) {
        for (int bi = b0; bi < b1; bi++) {
            const int64_t lo = start[bi], hi = start[(size_t)bi + 1];
            for (int64_t q = lo; q < hi; q++) {  // input order: a repeated (i,j) overwrites (last wins)
                const int64_t e = perm[(size_t)q];
                const size_t blk = (size_t)ptrow[bi] + (size_t)slot[(size_t)q];
                indcol[blk] = jcol[e] / 4;
                coef[16 * blk + 4 * (size_t)(irow[e] % 4) + (size_t)(jcol[e] % 4)] = val[e];
            }
        }
    });
    return total;
}

// Reads a Matrix Market coordinate file the way the reference's drivers do.  On success *irow, *jcol, *val are
// malloc'ed arrays of *nnz entries (0-based indices) released with nsk_mtx_free.
NSK_API int nsk_mtx_read(const char *path, int *nrow, int64_t *nnz, int **irow, int **jcol, double **val)
{
    if (!path || !nrow || !nnz || !irow || !jcol || !val) return NSK_ERR_INVALID;
    *irow = *jcol = nullptr;
    *val = nullptr;
    FILE *fp = fopen(path, "r");
    if (!fp) return NSK_ERR_INVALID;  // the reference prints "fail to open" and returns 1 (mpk/SpM2V.cpp:820-824)
    char buf[1024];
    int a = 0, b = 0, c = 0;
    bool ok = fgets(buf, sizeof buf, fp) != nullptr;  // banner: skipped unconditionally
    while (ok) {
        if (!fgets(buf, sizeof buf, fp)) { ok = false; break; }
        if (buf[0] != '%') {
            ok = sscanf(buf, "%d %d %d", &a, &b, &c) == 3;
            break;
        }
    }
    if (!ok || a < 0 || c < 0) { fclose(fp); return NSK_ERR_INVALID; }
    const int64_t m = c;
    int *ir = (int *)malloc(sizeof(int) * (size_t)std::max<int64_t>(m, 1));
    int *jc = (int *)malloc(sizeof(int) * (size_t)std::max<int64_t>(m, 1));
    double *v = (double *)malloc(sizeof(double) * (size_t)std::max<int64_t>(m, 1));
    if (!ir || !jc || !v) { free(ir); free(jc); free(v); fclose(fp); return NSK_ERR_ALLOC; }
    // the rest of the file in one read, then a hand-rolled scan: fscanf("%d %d %f") costs ~1 us per entry
    const long here = ftell(fp);
    fseek(fp, 0, SEEK_END);
    const long end = ftell(fp);
    fseek(fp, here, SEEK_SET);
    std::vector<char> text((size_t)(end - here) + 1);
    const size_t got = fread(text.data(), 1, (size_t)(end - here), fp);
    text[got] = 0;
    fclose(fp);
    char *p = text.data();
    for (int64_t e = 0; e < m; e++) {
        char *q;
        const long i = strtol(p, &q, 10);
        if (q == p) { free(ir); free(jc); free(v); return NSK_ERR_INVALID; }
        p = q;
        const long j = strtol(p, &q, 10);
        if (q == p) { free(ir); free(jc); free(v); return NSK_ERR_INVALID; }
        p = q;
        const float f = strtof(p, &q);  // "%f" into a float: values are fp32-rounded, then widened
        if (q == p) { free(ir); free(jc); free(v); return NSK_ERR_INVALID; }
        p = q;
        ir[e] = (int)i - 1;
        jc[e] = (int)j - 1;
        v[e] = (double)f;
    }
    *nrow = a;
    *nnz = m;
    *irow = ir;
    *jcol = jc;
    *val = v;
    return NSK_OK;
}

NSK_API void nsk_mtx_free(int *irow, int *jcol, double *val)
{
    free(irow);
    free(jcol);
    free(val);
}
