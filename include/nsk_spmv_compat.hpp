// nsk_spmv_compat.hpp -- C++ declarations that are link-compatible with the reference's mpk/SpMV.h.
//
// The reference's public interface for this path is a set of C++ free functions over two STL-based
// structs (reference mpk/SpMV.h:18-33 containers, :55-64 kernels; mpk/SpM2V.cpp:5,80,137,279 fused k=2).
// They are C++-mangled and take std::vector by reference, so they are not a C ABI; the drop-in is
// therefore two layers:
//
//     reference driver (mpk/2SpMV.cpp, mpk/SpM2V.cpp ...)      -- unchanged, includes ITS OWN SpMV.h
//        |  links against
//     libnsk_spmvshim.so  (navierstokes_b200/csrc/shim_spmv.cpp) -- same mangled symbols as mpk/SpMV.cpp
//        |  calls
//     libnsk.so           (include/nsk.h)                         -- the C ABI, CUDA kernels behind it
//
// The struct definitions below restate the reference's field order and types exactly -- that IS the
// binary contract -- and nothing else of the header.  Code that already includes the reference's
// SpMV.h must not include this file as well (same names).
#ifndef NSK_SPMV_COMPAT_HPP
#define NSK_SPMV_COMPAT_HPP

#include <vector>

struct csrmatrix {              // reference mpk/SpMV.h:18-24
    int n, nnz;
    std::vector<int> ptrow;     // n + 1
    std::vector<int> indcol;    // nnz, 0-based
    std::vector<double> coef;   // nnz
};

struct bcsr4x4_matrix {         // reference mpk/SpMV.h:26-33
    int nrows;                  // block rows (= n / 4)
    int nblocks;                // left at 0 by the reference's builder (mpk/utils.cpp:78)
    std::vector<int> ptrow;     // nrows + 1
    std::vector<int> indcol;    // block columns, first-appearance order
    std::vector<double> coef;   // 16 per block, row-major
};

// y = A x.  Synchronous, outputs fully overwritten, x untouched -- the reference's contract.
// Arithmetic: SpMV_CSR -> separately rounded multiply-add chain (the x87 original is not reproducible on a
// GPU); _OPT, _FMA -> fma chain, bit-identical to the reference; _AVX2 -> reassociated fast mode.
void SpMV_CSR(double *y, double *x, csrmatrix &A);        // reference mpk/SpMV.cpp:6
void SpMV_CSR_OPT(double *y, double *x, csrmatrix &A);    // reference mpk/SpMV.cpp:23
void SpMV_CSR_FMA(double *y, double *x, csrmatrix &A);    // reference mpk/SpMV.cpp:41
void SpMV_CSR_AVX2(double *y, double *x, csrmatrix &A);   // reference mpk/SpMV.cpp:59

void SpMV_BCSR(double *y, const double *x, const bcsr4x4_matrix &A);       // reference mpk/SpMV.cpp:90
void SpMV_BCSR_OPT(double *y, const double *x, const bcsr4x4_matrix &A);   // reference mpk/SpMV.cpp:121
void SpMV_BCSR_FMA(double *y, const double *x, const bcsr4x4_matrix &A);   // reference mpk/SpMV.cpp:154
void SpMV_BCSR_AVX2(double *y, const double *x, const bcsr4x4_matrix &A);  // reference mpk/SpMV.cpp:181

// Fused z = A(Ax), y = Ax.  The first-touch schedule argument is accepted and ignored (the GPU plan is
// built when the operator is first seen); Generate1stlayer still fills it the reference's way so that
// code which inspects it keeps working.
void Generate1stlayer(std::vector<int> &ptrowend1, csrmatrix &A);                                        // mpk/SpM2V.cpp:5
void SpM2V_CSR(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);             // mpk/SpM2V.cpp:80
void SpM2V_CSR_OPT(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);         // mpk/SpM2V.cpp:137
void SpM2V_CSR_AVX2(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);        // mpk/SpM2V.cpp:279

// k = 2, 3, 4 with the signatures of mpk/SpMVmulti0.cpp (:44, :65, :132, :191); the nested first-touch schedules are
// accepted and ignored (Generate2ndlayer / Generate3rdlayer still fill them the reference's way).  These are the
// x87 / no-fma flavours in the reference: multiply-add chain here.
void SpMV(double *y, double *x, csrmatrix &A);                                                              // :259
void Generate2ndlayer(std::vector<std::vector<int> > &ptrowend2, csrmatrix &A, std::vector<int> &ptrowend1);    // :106
void Generate3rdlayer(std::vector<std::vector<std::vector<int> > > &ptrowend3, csrmatrix &A,                   // :157
                      std::vector<int> &ptrowend1, std::vector<std::vector<int> > &ptrowend2);
void SpM2V0(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);
void SpM2V(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1);
void SpM3V(double *w, double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1,
           std::vector<std::vector<int> > &ptrowend2);
void SpM4V(double *v, double *w, double *z, double *y, double *x, csrmatrix &A, std::vector<int> &ptrowend1,
           std::vector<std::vector<int> > &ptrowend2, std::vector<std::vector<std::vector<int> > > &ptrowend3);

// Fused A^2 x on the block operator (mpk/SpM2V.cpp:28, :376, :475, :567, :675).
void Generate1stlayer_BCSR4(std::vector<int> &ptrowendB, const bcsr4x4_matrix &A);
void SpM2V_BCSR(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void SpM2V_BCSR_OPT(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void SpM2V_BCSR_FMA(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);
void SpM2V_BCSR_AVX2(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &ptrowendB);

// Shim housekeeping (not in the reference): drop every cached device operator / the GPU context.
extern "C" void nsk_shim_reset(void);

#endif
