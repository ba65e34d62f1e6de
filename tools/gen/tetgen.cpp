// tetgen.cpp -- fast generator of BASELINE config 4: the P1 stiffness matrix (vol * grad(phi_i).grad(phi_j), the formula
// of the reference's src/integration.c:19-57,231-236) on the Kuhn 6-tetrahedra split of an m^3-cell cube with jittered
// interior nodes, + 1e-3 * diag, assembled row by row in a RANDOMLY PERMUTED node numbering (what an unstructured mesh
// file would give), ready for RCM (nsk_rcm).  Same operator family as navierstokes_b200/matgen.py tet_p1_laplacian, whose
// numpy / scipy assembly takes minutes per million nodes; this one is threaded and allocation-free per row.
// Benchmark tooling, not part of the product library.
// Build: g++ -O3 -std=c++17 -shared -fPIC -pthread -o tools/bin/libtetgen.so tools/gen/tetgen.cpp
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <thread>
#include <vector>

namespace {
inline uint64_t splitmix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double unit(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)

const int KUHN[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};

struct Gen {
    int m, g;
    double jitter;
    uint64_t seed;
    void coord(int i, int j, int k, double *p) const
    {
        p[0] = (double)i / m; p[1] = (double)j / m; p[2] = (double)k / m;
        if (jitter > 0 && i > 0 && i < m && j > 0 && j < m && k > 0 && k < m) {
            const uint64_t id = ((uint64_t)i * g + j) * g + k;
            for (int d = 0; d < 3; d++) p[d] += (2.0 * unit(splitmix(seed * 1000003ull + id * 3 + d)) - 1.0) * jitter / m;
        }
    }
};
}  // namespace

// Number of nodes: (m + 1)^3.  perm_seed < 0: natural numbering.  Two calls: sizes first (ptrow), then the entries.
extern "C" __attribute__((visibility("default")))
int64_t tetgen_rows(int m, double jitter, long long jitter_seed, long long perm_seed, int *ptrow, int *indcol, double *coef,
                    int nthreads)
{
    const int g = m + 1;
    const int64_t n64 = (int64_t)g * g * g;
    if (n64 >= 2147483647) return -1;
    const int n = (int)n64;
    Gen G{m, g, jitter, (uint64_t)jitter_seed};
    // new -> old and old -> new numbering
    std::vector<int> n2o(n), o2n(n);
    std::iota(n2o.begin(), n2o.end(), 0);
    if (perm_seed >= 0) {
        uint64_t s = (uint64_t)perm_seed * 0x2545F4914F6CDD1Dull + 7;
        for (int i = n - 1; i > 0; i--) {
            s = splitmix(s);
            const int j = (int)(s % (uint64_t)(i + 1));
            std::swap(n2o[i], n2o[j]);
        }
    }
    for (int i = 0; i < n; i++) o2n[n2o[i]] = i;
    const bool fill = indcol != nullptr && coef != nullptr;
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int tid) {
        int cols[32];
        double vals[32];
        for (int r = tid; r < n; r += nthreads) {  // interleaved rows: the permutation already spreads the work
            const int old = n2o[r];
            const int k = old % g, j = (old / g) % g, i = old / (g * g);
            int cnt = 0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    for (int c = 0; c < 2; c++) {
                        const int ci = i - a, cj = j - b, ck = k - c;  // cell whose corner (a, b, c) is this node
                        if (ci < 0 || ci >= m || cj < 0 || cj >= m || ck < 0 || ck >= m) continue;
                        const int me = (a << 2) | (b << 1) | c;
                        for (int t = 0; t < 6; t++) {
                            int li = -1;
                            for (int v = 0; v < 4; v++)
                                if (KUHN[t][v] == me) li = v;
                            if (li < 0) continue;
                            if (!fill) {  // pattern only: the tet's four nodes
                                for (int v = 0; v < 4; v++) {
                                    const int cb = KUHN[t][v];
                                    const int id = o2n[((ci + ((cb >> 2) & 1)) * g + (cj + ((cb >> 1) & 1))) * g + (ck + (cb & 1))];
                                    int q = 0;
                                    while (q < cnt && cols[q] != id) q++;
                                    if (q == cnt) cols[cnt++] = id;
                                }
                                continue;
                            }
                            double p[4][3];
                            int ids[4];
                            for (int v = 0; v < 4; v++) {
                                const int cb = KUHN[t][v];
                                const int ni = ci + ((cb >> 2) & 1), nj = cj + ((cb >> 1) & 1), nk = ck + (cb & 1);
                                G.coord(ni, nj, nk, p[v]);
                                ids[v] = o2n[(ni * g + nj) * g + nk];
                            }
                            double d[3][3];
                            for (int e = 0; e < 3; e++)
                                for (int x = 0; x < 3; x++) d[e][x] = p[e + 1][x] - p[0][x];
                            const double det = d[0][0] * (d[1][1] * d[2][2] - d[1][2] * d[2][1]) -
                                               d[0][1] * (d[1][0] * d[2][2] - d[1][2] * d[2][0]) +
                                               d[0][2] * (d[1][0] * d[2][1] - d[1][1] * d[2][0]);
                            const double vol = std::fabs(det) / 6.0, id_ = 1.0 / det;
                            // gradients of phi_1..3 = columns of inverse(d); phi_0 = minus their sum
                            double gr[4][3];
                            gr[1][0] = (d[1][1] * d[2][2] - d[1][2] * d[2][1]) * id_;
                            gr[1][1] = (d[1][2] * d[2][0] - d[1][0] * d[2][2]) * id_;
                            gr[1][2] = (d[1][0] * d[2][1] - d[1][1] * d[2][0]) * id_;
                            gr[2][0] = (d[0][2] * d[2][1] - d[0][1] * d[2][2]) * id_;
                            gr[2][1] = (d[0][0] * d[2][2] - d[0][2] * d[2][0]) * id_;
                            gr[2][2] = (d[0][1] * d[2][0] - d[0][0] * d[2][1]) * id_;
                            gr[3][0] = (d[0][1] * d[1][2] - d[0][2] * d[1][1]) * id_;
                            gr[3][1] = (d[0][2] * d[1][0] - d[0][0] * d[1][2]) * id_;
                            gr[3][2] = (d[0][0] * d[1][1] - d[0][1] * d[1][0]) * id_;
                            for (int x = 0; x < 3; x++) gr[0][x] = -(gr[1][x] + gr[2][x] + gr[3][x]);
                            for (int v = 0; v < 4; v++) {
                                const double ke = vol * (gr[li][0] * gr[v][0] + gr[li][1] * gr[v][1] + gr[li][2] * gr[v][2]);
                                int q = 0;
                                while (q < cnt && cols[q] != ids[v]) q++;
                                if (q == cnt) { cols[cnt] = ids[v]; vals[cnt] = 0.0; cnt++; }
                                vals[q] += ke;
                            }
                        }
                    }
            if (!fill) {
                ptrow[r + 1] = cnt;
                continue;
            }
            // ascending columns, diagonal shifted by 1e-3 of itself
            int order[32];
            for (int q = 0; q < cnt; q++) order[q] = q;
            std::sort(order, order + cnt, [&](int x, int y) { return cols[x] < cols[y]; });
            const int base = ptrow[r];
            for (int q = 0; q < cnt; q++) {
                const int s = order[q];
                indcol[base + q] = cols[s];
                coef[base + q] = cols[s] == r ? vals[s] * 1.001 : vals[s];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (!fill) {
        ptrow[0] = 0;
        int64_t acc = 0;
        for (int r = 0; r < n; r++) {
            acc += ptrow[r + 1];
            if (acc >= 2147483647) return -2;
            ptrow[r + 1] = (int)acc;
        }
        return acc;
    }
    return ptrow[n];
}
