// ref_wrap_spmv.cpp -- C-linkage doorway into the UNMODIFIED reference objects.
//
// TEST INFRASTRUCTURE ONLY (see oracle/nsk_oracle.c header).  This file contains no reference
// code: it declares the reference's entry points (mpk/SpMV.h:37-66, mpk/SpM2V.cpp:5,28,80,137,
// mpk/2SpMV.cpp:3) and forwards raw pointers into them, so that Python tests (ctypes) and
// bench.py's `--impl reference` leg can call the real thing.  It is compiled by
// oracle/Makefile together with the reference sources *where they lie* under
// /root/reference/mpk into oracle/_ref/libnsref_spmv.so (git-ignored, travels with gpurun).
#include "SpMV.h"   // found with -I/root/reference/mpk at build time; never copied

// Defined in mpk/SpM2V.cpp (compiled with -Dmain=ref_spm2v_main) and mpk/2SpMV.cpp
// (compiled with -Dmain=ref_2spmv_main); they have no header in the reference.
void Generate1stlayer(std::vector<int> &ptrowend1, csrmatrix &A);
void Generate1stlayer_BCSR4(std::vector<int> &ptrowendB, const bcsr4x4_matrix &A);
void SpM2V_CSR(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);
void SpM2V_CSR_OPT(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);
void SpM2V_CSR_AVX2(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);
void SpM2V_BCSR(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void SpM2V_BCSR_OPT(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void SpM2V_BCSR_AVX2(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void orthogonalize(int nrow, const std::vector<double> &x, std::vector<double> &y, double alpha);
void orthonormalize_against_basis(int nrow, std::vector<std::vector<double>> &basis, std::vector<double> &y);

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
bcsr4x4_matrix make_bcsr(int nbrows, int nblk, const int *ptrow, const int *indcol,
                         const double *coef)
{
    bcsr4x4_matrix B;
    B.nrows = nbrows;
    B.nblocks = 0;
    B.ptrow.assign(ptrow, ptrow + nbrows + 1);
    B.indcol.assign(indcol, indcol + nblk);
    B.coef.assign(coef, coef + 16 * (size_t)nblk);
    return B;
}
}  // namespace

extern "C" {

// A persistent handle avoids re-copying a 1.4 GB matrix for every timed call.
void *ref_csr_new(int n, int nnz, const int *ptrow, const int *indcol, const double *coef)
{
    return new csrmatrix(make_csr(n, nnz, ptrow, indcol, coef));
}
void ref_csr_free(void *h) { delete static_cast<csrmatrix *>(h); }

// variant: 0 = SpMV_CSR (x87), 1 = _OPT, 2 = _FMA, 3 = _AVX2   (mpk/SpMV.h:55-58)
int ref_spmv_csr(void *h, int variant, const double *x, double *y)
{
    csrmatrix &A = *static_cast<csrmatrix *>(h);
    double *xx = const_cast<double *>(x);
    switch (variant) {
        case 0: SpMV_CSR(y, xx, A); return 0;
        case 1: SpMV_CSR_OPT(y, xx, A); return 0;
        case 2: SpMV_CSR_FMA(y, xx, A); return 0;
        case 3: SpMV_CSR_AVX2(y, xx, A); return 0;
    }
    return -1;
}

void ref_generate_1st_layer(void *h, int *ptrowend1)
{
    csrmatrix &A = *static_cast<csrmatrix *>(h);
    std::vector<int> pe;
    Generate1stlayer(pe, A);
    std::copy(pe.begin(), pe.end(), ptrowend1);
}

// variant: 0 = SpM2V_CSR (x87), 1 = SpM2V_CSR_OPT, 3 = SpM2V_CSR_AVX2 (mpk/SpM2V.cpp:80,137,278)
int ref_spm2v_csr(void *h, int variant, const int *ptrowend1, const double *x, double *y,
                  double *z)
{
    csrmatrix &A = *static_cast<csrmatrix *>(h);
    std::vector<int> pe(ptrowend1, ptrowend1 + A.nnz);
    double *xx = const_cast<double *>(x);
    switch (variant) {
        case 0: SpM2V_CSR(z, y, xx, A, pe); return 0;
        case 1: SpM2V_CSR_OPT(z, y, xx, A, pe); return 0;
        case 3: SpM2V_CSR_AVX2(z, y, xx, A, pe); return 0;
    }
    return -1;
}

// COO2CSR (mpk/utils.cpp:97-127).  Returns ptrow[nrow] (= entries kept after duplicate drop).
int ref_coo2csr(int nrow, int nnz, const int *irow, const int *jcol, const double *val,
                int *ptrow, int *indcol, double *coef)
{
    csrmatrix A;
    COO2CSR(A, nrow, nnz, const_cast<int *>(irow), const_cast<int *>(jcol),
            const_cast<double *>(val));
    std::copy(A.ptrow.begin(), A.ptrow.end(), ptrow);
    std::copy(A.indcol.begin(), A.indcol.end(), indcol);
    std::copy(A.coef.begin(), A.coef.end(), coef);
    return A.ptrow[nrow];
}

// generate_BCSR4 (mpk/utils.cpp:45-95).  Two-call protocol like the oracle's.
int ref_generate_bcsr4(int nrow, int nnz, const int *irow, const int *jcol, const double *val,
                       int *ptrow, int *indcol, double *coef)
{
    bcsr4x4_matrix B;
    std::vector<std::list<std::pair<int, std::array<double, 16>>>> block_rows((nrow + 3) / 4);
    generate_BCSR4(&block_rows[0], nrow, nnz, irow, jcol, val, B);
    int nblk = (int)B.indcol.size();
    if (indcol) {
        std::copy(B.ptrow.begin(), B.ptrow.end(), ptrow);
        std::copy(B.indcol.begin(), B.indcol.end(), indcol);
        std::copy(B.coef.begin(), B.coef.end(), coef);
    }
    return nblk;
}

// variant: 0 = SpMV_BCSR (x87), 1 = _OPT, 2 = _FMA, 3 = _AVX2   (mpk/SpMV.h:61-64)
int ref_spmv_bcsr4(int nbrows, int nblk, const int *ptrow, const int *indcol,
                   const double *coef, int variant, const double *x, double *y)
{
    bcsr4x4_matrix B = make_bcsr(nbrows, nblk, ptrow, indcol, coef);
    switch (variant) {
        case 0: SpMV_BCSR(y, x, B); return 0;
        case 1: SpMV_BCSR_OPT(y, x, B); return 0;
        case 2: SpMV_BCSR_FMA(y, x, B); return 0;
        case 3: SpMV_BCSR_AVX2(y, x, B); return 0;
    }
    return -1;
}

// Generate1stlayer_BCSR4 + SpM2V_BCSR{,_OPT,_AVX2} (mpk/SpM2V.cpp:28-46, 376, 475, 675).
// variant: 0 = x87, 1 = _OPT, 3 = _AVX2.  y,z have length 4*nbrows.
int ref_spm2v_bcsr4(int nbrows, int nblk, const int *ptrow, const int *indcol,
                    const double *coef, int variant, const double *x, double *y, double *z)
{
    bcsr4x4_matrix B = make_bcsr(nbrows, nblk, ptrow, indcol, coef);
    std::vector<int> pe;
    Generate1stlayer_BCSR4(pe, B);
    double *xx = const_cast<double *>(x);
    switch (variant) {
        case 0: SpM2V_BCSR(z, y, xx, B, pe); return 0;
        case 1: SpM2V_BCSR_OPT(z, y, xx, B, pe); return 0;
        case 3: SpM2V_BCSR_AVX2(z, y, xx, B, pe); return 0;
    }
    return -1;
}

double ref_norm2(int n, const double *x)
{
    std::vector<double> v(x, x + n);
    return norm2(v);
}

double ref_rel_error(int n, const double *ref, const double *test)
{
    std::vector<double> a(ref, ref + n), b(test, test + n);
    return rel_error(a, b);
}

void ref_orthogonalize(int n, const double *x, double *y, double alpha)
{
    std::vector<double> xv(x, x + n), yv(y, y + n);
    orthogonalize(n, xv, yv, alpha);
    std::copy(yv.begin(), yv.end(), y);
}

void ref_orthonormalize_against_basis(int n, int m, const double *basis, double *y)
{
    std::vector<std::vector<double>> B(m);
    for (int j = 0; j < m; j++) B[j].assign(basis + (size_t)j * n, basis + (size_t)(j + 1) * n);
    std::vector<double> yv(y, y + n);
    orthonormalize_against_basis(n, B, yv);
    std::copy(yv.begin(), yv.end(), y);
}

void ref_flush_cache(void) { flush_cache(); }

}  // extern "C"
