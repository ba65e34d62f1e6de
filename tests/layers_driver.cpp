// layers_driver.cpp -- CPU check of the shim's first-touch schedule builders against the reference's own.
//
// Loads libnsk_spmvshim.so and oracle/_ref/libnsref_multi0.so side by side (RTLD_LOCAL, so the identically
// named symbols do not collide), builds random square CSR operators, and compares Generate{1st,2nd,3rd}layer
// entry for entry.  Compiled and run by tests/test_abi.py; prints "OK <cases>" or the first difference.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
#include "nsk_spmv_compat.hpp"

typedef std::vector<int> L1;
typedef std::vector<L1> L2;
typedef std::vector<L2> L3;
typedef void (*gen1_t)(L1 &, csrmatrix &);
typedef void (*gen2_t)(L2 &, csrmatrix &, L1 &);
typedef void (*gen3_t)(L3 &, csrmatrix &, L1 &, L2 &);

static void *need(void *h, const char *name)
{
    void *p = dlsym(h, name);
    if (!p) { std::fprintf(stderr, "missing %s: %s\n", name, dlerror()); std::exit(2); }
    return p;
}

static csrmatrix random_csr(int n, int max_row, bool sorted, unsigned seed)
{
    std::mt19937 rng(seed);
    csrmatrix A;
    A.n = n;
    A.ptrow.push_back(0);
    for (int i = 0; i < n; i++) {
        const int len = max_row > 0 ? (int)(rng() % (unsigned)(max_row + 1)) : 0;
        std::vector<int> cols;
        for (int t = 0; t < len; t++) {
            const int c = (int)(rng() % (unsigned)n);
            if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c);
        }
        if (sorted) std::sort(cols.begin(), cols.end());
        for (int c : cols) { A.indcol.push_back(c); A.coef.push_back(1.0 + (double)(rng() % 7)); }
        A.ptrow.push_back((int)A.indcol.size());
    }
    A.nnz = (int)A.indcol.size();
    return A;
}

int main(int argc, char **argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s shim.so ref_multi0.so\n", argv[0]); return 2; }
    void *hs = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    void *hr = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!hs || !hr) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    const char *n1 = "_Z16Generate1stlayerRSt6vectorIiSaIiEER9csrmatrix";
    const char *n2 = "_Z16Generate2ndlayerRSt6vectorIS_IiSaIiEESaIS1_EER9csrmatrixRS1_";
    const char *n3 = "_Z16Generate3rdlayerRSt6vectorIS_IS_IiSaIiEESaIS1_EESaIS3_EER9csrmatrixRS1_RS3_";
    gen1_t s1 = (gen1_t)need(hs, n1), r1 = (gen1_t)need(hr, n1);
    gen2_t s2 = (gen2_t)need(hs, n2), r2 = (gen2_t)need(hr, n2);
    gen3_t s3 = (gen3_t)need(hs, n3), r3 = (gen3_t)need(hr, n3);

    int cases = 0;
    const int sizes[] = {1, 2, 7, 40, 150};
    for (int n : sizes)
        for (int max_row = 0; max_row <= 6; max_row += 2)
            for (int sorted = 0; sorted < 2; sorted++) {
                csrmatrix A = random_csr(n, std::min(max_row, n), sorted != 0, 1000u * n + 10u * max_row + sorted);
                if (A.nnz == 0) continue;   // the reference's builder indexes an empty vector on an empty operator
                L1 a1, b1; L2 a2, b2; L3 a3, b3;
                s1(a1, A); r1(b1, A);
                s2(a2, A, a1); r2(b2, A, b1);
                s3(a3, A, a1, a2); r3(b3, A, b1, b2);
                if (a1 != b1 || a2 != b2 || a3 != b3) {
                    std::printf("MISMATCH n=%d max_row=%d sorted=%d: depth1 %d depth2 %d depth3 %d\n", n, max_row, sorted,
                                (int)(a1 == b1), (int)(a2 == b2), (int)(a3 == b3));
                    return 1;
                }
                cases++;
            }
    std::printf("OK %d\n", cases);
    return 0;
}
