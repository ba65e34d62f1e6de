// ref_wrap_multi0.cpp -- C-linkage doorway into the reference's matrix-powers seed file.
//
// TEST INFRASTRUCTURE ONLY.  mpk/SpMVmulti0.cpp is self-contained (its own csrmatrix, COO2CSR,
// SpMV, SpM2V/3V/4V, Generate{1st,2nd,3rd}layer and a main); it is pulled in *from where it
// lies* under /root/reference at build time with its main renamed, so this file holds no
// reference code.  Built by oracle/Makefile into oracle/_ref/libnsref_multi0.so.
#define main ref_multi0_main
#include "SpMVmulti0.cpp"   // -I/root/reference/mpk ; never copied into this repository
#undef main

namespace {
csrmatrix make_csr(int n, int nnz, const int *ptrow, const int *indcol, const double *coef)
{
    csrmatrix A;
    A.n = n;
    A.nnz = nnz;
    A.ptrow.assign(ptrow, ptrow + n + 1);
    A.indcol.assign(indcol, indcol + nnz);
    A.coef.assign(coef, coef + nnz);
    return A;
}
}  // namespace

extern "C" {

// Plain SpMV of the seed file (mpk/SpMVmulti0.cpp:259-270, x87 arithmetic).
void ref0_spmv(int n, int nnz, const int *ptrow, const int *indcol, const double *coef,
               const double *x, double *y)
{
    csrmatrix A = make_csr(n, nnz, ptrow, indcol, coef);
    SpMV(y, const_cast<double *>(x), A);
}

// Fused A^k x for k = 2, 3, 4 through the reference's own schedule builders
// (Generate1stlayer :22, Generate2ndlayer :106, Generate3rdlayer :157) and kernels
// (SpM2V0 :44, SpM3V :132, SpM4V :191).  out = [y | z | w | v], each n long.
// The nested-vector schedules are O(nnz * row^2) in memory: toy sizes only.
int ref0_spmkv(int n, int nnz, const int *ptrow, const int *indcol, const double *coef,
               int depth, const double *x, double *out)
{
    csrmatrix A = make_csr(n, nnz, ptrow, indcol, coef);
    double *xx = const_cast<double *>(x);
    double *y = out, *z = out + n, *w = out + 2 * (size_t)n, *v = out + 3 * (size_t)n;
    std::vector<int> pe1;
    Generate1stlayer(pe1, A);
    if (depth == 2) { SpM2V0(z, y, xx, A, pe1); return 0; }
    std::vector<std::vector<int> > pe2;
    Generate2ndlayer(pe2, A, pe1);
    if (depth == 3) { SpM3V(w, z, y, xx, A, pe1, pe2); return 0; }
    std::vector<std::vector<std::vector<int> > > pe3;
    Generate3rdlayer(pe3, A, pe1, pe2);
    if (depth == 4) { SpM4V(v, w, z, y, xx, A, pe1, pe2, pe3); return 0; }
    return -1;
}

}  // extern "C"
