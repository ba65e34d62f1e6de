// Exercises the C++ drop-in shim (libnsk_spmvshim.so) through the reference's own function names and STL-based
// containers (include/nsk_spmv_compat.hpp restates mpk/SpMV.h:18-33), and compares every output bit for bit with a
// plain host loop of the same per-row chain.  Built and run by tests/test_shim_gpu.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "nsk_spmv_compat.hpp"

static void host_spmv(const csrmatrix &A, const std::vector<double> &x, std::vector<double> &y, bool use_fma)
{
    for (int i = 0; i < A.n; i++) {
        double acc = 0.0;
        for (int j = A.ptrow[i]; j < A.ptrow[i + 1]; j++) {
            if (use_fma) acc = std::fma(A.coef[j], x[A.indcol[j]], acc);
            else { volatile double p = A.coef[j] * x[A.indcol[j]]; acc = acc + p; }
        }
        y[i] = acc;
    }
}

static bool same(const std::vector<double> &a, const std::vector<double> &b)
{
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0;
}

int main()
{
    // 3D 7-point Laplacian 20 x 12 x 10 (n = 2400, a multiple of 4 for the block format)
    const int nx = 20, ny = 12, nz = 10, n = nx * ny * nz;
    csrmatrix A;
    A.n = n;
    A.ptrow.push_back(0);
    for (int i = 0; i < n; i++) {
        const int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
        auto add = [&](bool ok, int off, double v) { if (ok) { A.indcol.push_back(i + off); A.coef.push_back(v); } };
        add(iz > 0, -nx * ny, -1.0); add(iy > 0, -nx, -1.0); add(ix > 0, -1, -1.0);
        add(true, 0, 6.0 + 0.001 * i);
        add(ix < nx - 1, 1, -1.0); add(iy < ny - 1, nx, -1.0); add(iz < nz - 1, nx * ny, -1.0);
        A.ptrow.push_back((int)A.indcol.size());
    }
    A.nnz = (int)A.indcol.size();
    std::vector<double> x(n);
    for (int i = 0; i < n; i++) x[i] = std::sin(0.37 * i) + 0.25;

    int bad = 0;
    std::vector<double> y(n), z(n), w(n), v(n), r1(n), r2(n), r3(n), r4(n);
    std::vector<int> pe1;
    std::vector<std::vector<int> > pe2;
    std::vector<std::vector<std::vector<int> > > pe3;

    // fma flavours
    host_spmv(A, x, r1, true); host_spmv(A, r1, r2, true);
    SpMV_CSR_FMA(y.data(), x.data(), A);
    bad += !same(y, r1);
    Generate1stlayer(pe1, A);
    bad += (int)pe1.size() != A.nnz;
    SpM2V_CSR_OPT(z.data(), y.data(), x.data(), A, pe1);
    bad += !same(y, r1) + !same(z, r2);
    // multiply-add flavours (the reference compiles these no-fma): SpM2V0 / SpM2V / SpM3V / SpM4V of SpMVmulti0.cpp
    host_spmv(A, x, r1, false); host_spmv(A, r1, r2, false); host_spmv(A, r2, r3, false); host_spmv(A, r3, r4, false);
    SpM2V(z.data(), y.data(), x.data(), A, pe1);
    bad += !same(y, r1) + !same(z, r2);
    SpM3V(w.data(), z.data(), y.data(), x.data(), A, pe1, pe2);
    bad += !same(y, r1) + !same(z, r2) + !same(w, r3);
    SpM4V(v.data(), w.data(), z.data(), y.data(), x.data(), A, pe1, pe2, pe3);
    bad += !same(y, r1) + !same(z, r2) + !same(w, r3) + !same(v, r4);

    // block format: one diagonal-ish block structure built directly (block row bi touches bi-1, bi, bi+1)
    bcsr4x4_matrix B;
    B.nrows = n / 4;
    B.nblocks = 0;
    B.ptrow.push_back(0);
    for (int bi = 0; bi < B.nrows; bi++) {
        for (int bj = bi - 1; bj <= bi + 1; bj++) {
            if (bj < 0 || bj >= B.nrows) continue;
            B.indcol.push_back(bj);
            for (int e = 0; e < 16; e++) B.coef.push_back(std::cos(0.01 * (bi * 48 + (bj - bi + 1) * 16 + e)));
        }
        B.ptrow.push_back((int)B.indcol.size());
    }
    auto host_bcsr = [&](const std::vector<double> &in, std::vector<double> &out) {
        for (int bi = 0; bi < B.nrows; bi++)
            for (int i = 0; i < 4; i++) {
                double acc = 0.0;
                for (int m = B.ptrow[bi]; m < B.ptrow[bi + 1]; m++)
                    for (int j = 0; j < 4; j++) acc = std::fma(B.coef[16 * m + 4 * i + j], in[4 * B.indcol[m] + j], acc);
                out[4 * bi + i] = acc;
            }
    };
    host_bcsr(x, r1); host_bcsr(r1, r2);
    SpMV_BCSR_FMA(y.data(), x.data(), B);
    bad += !same(y, r1);
    std::vector<int> peB;
    Generate1stlayer_BCSR4(peB, B);
    bad += peB.size() != B.indcol.size();
    SpM2V_BCSR_OPT(z.data(), y.data(), x.data(), B, peB);
    bad += !same(y, r1) + !same(z, r2);
    SpM2V_BCSR_AVX2(z.data(), y.data(), x.data(), B, peB);
    bad += !same(y, r1) + !same(z, r2);

    // the caller updates the coefficients IN PLACE (same addresses, same pattern: a new Jacobian every Newton step) and
    // calls again without telling anyone -- the reference reads the live arrays, so must the shim
    for (size_t j = 0; j < A.coef.size(); j++) A.coef[j] = A.coef[j] * 1.25 + 0.5;
    host_spmv(A, x, r1, true); host_spmv(A, r1, r2, true);
    SpMV_CSR_FMA(y.data(), x.data(), A);
    bad += !same(y, r1);
    SpM2V_CSR_OPT(z.data(), y.data(), x.data(), A, pe1);
    bad += !same(y, r1) + !same(z, r2);
    for (size_t j = 0; j < B.coef.size(); j++) B.coef[j] = -B.coef[j];
    host_bcsr(x, r1);
    SpMV_BCSR_FMA(y.data(), x.data(), B);
    bad += !same(y, r1);

    nsk_shim_reset();
    std::printf("shim_driver: %s (%d mismatching outputs)\n", bad ? "FAILED" : "OK", bad);
    return bad ? 1 : 0;
}
