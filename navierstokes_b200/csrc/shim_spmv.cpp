// shim_spmv.cpp -- the reference's SpMV.h entry points, implemented on the C ABI (include/nsk.h).
//
// Built into libnsk_spmvshim.so.  A reference driver (mpk/2SpMV.cpp ...) linked against this library
// instead of mpk/SpMV.cpp runs unchanged on the GPU (INTEGRATION.md shows the link line).
//
// Ownership follows the reference: the caller owns the matrix (std::vector storage) and
// the host vectors; kernels are synchronous and allocate nothing the caller can see.  The device copy
// of an operator is cached, keyed on the identity of the caller's arrays, so the upload happens once
// per matrix like the reference's one-time COO2CSR.  Errors: the reference's functions are `void` and
// cannot fail; here a failure (no GPU, out of memory) prints the library's message and aborts -- there
// is NO CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "nsk.h"
#include "nsk_spmv_compat.hpp"

namespace {

std::mutex g_mu;
nsk_ctx_t g_ctx = nullptr;

[[noreturn]] void die(const char *what, int status)
{
    std::fprintf(stderr, "nsk shim: %s failed: %s (%s)\n", what, nsk_strerror(status), nsk_last_error(g_ctx));
    std::abort();
}

nsk_ctx_t ctx()
{
    if (!g_ctx) {
        const char *dev = std::getenv("NSK_DEVICE");
        int s = nsk_ctx_create(dev ? std::atoi(dev) : 0, &g_ctx);
        if (s != NSK_OK) die("nsk_ctx_create", s);
    }
    return g_ctx;
}

typedef std::tuple<const void *, const void *, const void *, int, long long> Key;
// The reference's kernels read the caller's live arrays on every call; the shim keeps a device copy keyed on the
// arrays' addresses and sizes.  A caller may update coef in place (new Jacobian values, same pattern) or free a matrix
// and get the same addresses back, so every call re-validates the entry with a fingerprint of the host arrays and
// re-uploads the operator when it changed.  NSK_SHIM_VALIDATE = "sample" (default: 4096 windows of 64 bytes spread over
// each array, plus both ends -- catches any wholesale update), "full" (every byte) or "off".
template <class H>
struct Entry {
    H handle = nullptr;
    unsigned long long print = 0;
};
std::map<Key, Entry<nsk_csr_t>> g_csr;
std::map<Key, Entry<nsk_bcsr4_t>> g_bcsr;

int validate_mode()
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = std::getenv("NSK_SHIM_VALIDATE");
        mode = !e ? 1 : !std::strcmp(e, "off") ? 0 : !std::strcmp(e, "full") ? 2 : 1;
    }
    return mode;
}

unsigned long long mix(unsigned long long h, const void *p, size_t bytes)
{
    const unsigned char *b = static_cast<const unsigned char *>(p);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        unsigned long long w;
        std::memcpy(&w, b + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    for (; i < bytes; i++) h = (h ^ b[i]) * 0x100000001B3ull;
    return h;
}

unsigned long long fingerprint_array(unsigned long long h, const void *p, size_t bytes)
{
    const int mode = validate_mode();
    if (mode == 0 || bytes == 0) return h;
    if (mode == 2 || bytes <= 4096 * 64 * 2) return mix(h, p, bytes);
    const unsigned char *b = static_cast<const unsigned char *>(p);
    const size_t windows = 4096, step = (bytes - 64) / (windows - 1);
    for (size_t w = 0; w < windows; w++) h = mix(h, b + w * step, 64);
    return mix(h, b + bytes - 64, 64);
}

unsigned long long fingerprint(const void *ptrow, size_t pb, const void *indcol, size_t ib, const void *coef, size_t cb)
{
    unsigned long long h = 0xcbf29ce484222325ull;
    h = fingerprint_array(h, ptrow, pb);
    h = fingerprint_array(h, indcol, ib);
    return fingerprint_array(h, coef, cb);
}

nsk_csr_t device_csr(csrmatrix &A)
{
    // the reference leaves A.nnz at the COO count even when duplicates were dropped (mpk/utils.cpp:100):
    // the true count is ptrow[n]
    const long long nnz = A.ptrow.empty() ? 0 : (long long)A.ptrow[A.n];
    Key k(A.ptrow.data(), A.indcol.data(), A.coef.data(), A.n, nnz);
    const unsigned long long fp = fingerprint(A.ptrow.data(), sizeof(int) * ((size_t)A.n + 1), A.indcol.data(),
                                              sizeof(int) * (size_t)nnz, A.coef.data(), sizeof(double) * (size_t)nnz);
    auto it = g_csr.find(k);
    if (it != g_csr.end()) {
        if (it->second.print == fp) return it->second.handle;
        nsk_csr_destroy(it->second.handle);  // same addresses, different contents: the device copy is stale
        g_csr.erase(it);
    }
    nsk_csr_t h = nullptr;
    int s = nsk_csr_create(ctx(), A.n, A.n, nnz, A.ptrow.data(), A.indcol.data(), A.coef.data(), &h);
    if (s != NSK_OK) die("nsk_csr_create", s);
    g_csr[k] = Entry<nsk_csr_t>{h, fp};
    return h;
}

nsk_bcsr4_t device_bcsr(const bcsr4x4_matrix &B)
{
    const long long nblk = (long long)B.indcol.size();
    Key k(B.ptrow.data(), B.indcol.data(), B.coef.data(), B.nrows, nblk);
    const unsigned long long fp = fingerprint(B.ptrow.data(), sizeof(int) * ((size_t)B.nrows + 1), B.indcol.data(),
                                              sizeof(int) * (size_t)nblk, B.coef.data(), sizeof(double) * B.coef.size());
    auto it = g_bcsr.find(k);
    if (it != g_bcsr.end()) {
        if (it->second.print == fp) return it->second.handle;
        nsk_bcsr4_destroy(it->second.handle);
        g_bcsr.erase(it);
    }
    nsk_bcsr4_t h = nullptr;
    int s = nsk_bcsr4_create(ctx(), B.nrows, nblk, B.ptrow.data(), B.indcol.data(), B.coef.data(), &h);
    if (s != NSK_OK) die("nsk_bcsr4_create", s);
    g_bcsr[k] = Entry<nsk_bcsr4_t>{h, fp};
    return h;
}

void spmv(double *y, const double *x, csrmatrix &A, nsk_mode mode)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int s = nsk_spmv(device_csr(A), x, y, mode, NSK_HOST);
    if (s != NSK_OK) die("nsk_spmv", s);
}

void spmv_b(double *y, const double *x, const bcsr4x4_matrix &B, nsk_mode mode)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int s = nsk_spmv_bcsr4(device_bcsr(B), x, y, mode, NSK_HOST);
    if (s != NSK_OK) die("nsk_spmv_bcsr4", s);
}

void spm2v(double *z, double *y, const double *x, csrmatrix &A, nsk_mode mode)
{
    std::lock_guard<std::mutex> lk(g_mu);
    double *levels[2] = {y, z};
    int s = nsk_mpk(device_csr(A), 2, x, levels, mode, NSK_HOST);
    if (s != NSK_OK) die("nsk_mpk", s);
}

void spmkv(int k, double *const *levels, const double *x, csrmatrix &A, nsk_mode mode)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int s = nsk_mpk(device_csr(A), k, x, levels, mode, NSK_HOST);
    if (s != NSK_OK) die("nsk_mpk", s);
}

void spm2v_b(double *z, double *y, const double *x, const bcsr4x4_matrix &B, nsk_mode mode)
{
    std::lock_guard<std::mutex> lk(g_mu);
    nsk_bcsr4_t h = device_bcsr(B);
    double *levels[2] = {y, z};
    int s = nsk_bcsr4_mpk(h, 2, x, levels, mode, NSK_HOST);  // one call: x in once, both levels out
    if (s != NSK_OK) die("nsk_bcsr4_mpk", s);
}

}  // namespace

void SpMV_CSR(double *y, double *x, csrmatrix &A) { spmv(y, x, A, NSK_EXACT_MULADD); }
void SpMV_CSR_OPT(double *y, double *x, csrmatrix &A) { spmv(y, x, A, NSK_EXACT_FMA); }
void SpMV_CSR_FMA(double *y, double *x, csrmatrix &A) { spmv(y, x, A, NSK_EXACT_FMA); }
void SpMV_CSR_AVX2(double *y, double *x, csrmatrix &A) { spmv(y, x, A, NSK_FAST); }

void SpMV_BCSR(double *y, const double *x, const bcsr4x4_matrix &A) { spmv_b(y, x, A, NSK_EXACT_MULADD); }
void SpMV_BCSR_OPT(double *y, const double *x, const bcsr4x4_matrix &A) { spmv_b(y, x, A, NSK_EXACT_FMA); }
void SpMV_BCSR_FMA(double *y, const double *x, const bcsr4x4_matrix &A) { spmv_b(y, x, A, NSK_EXACT_FMA); }
void SpMV_BCSR_AVX2(double *y, const double *x, const bcsr4x4_matrix &A) { spmv_b(y, x, A, NSK_FAST); }

// Same contract as the reference's schedule builder: entry ia gets the end of row indcol[ia] the first
// time that column is met in row-major order, its beginning afterwards.  The GPU kernels do not read it.
void Generate1stlayer(std::vector<int> &ptrowend1, csrmatrix &A)
{
    const int n = A.n;
    const int nnz = A.ptrow.empty() ? 0 : A.ptrow[n];
    std::vector<char> seen(n > 0 ? n : 1, 0);
    ptrowend1.assign(A.nnz > nnz ? A.nnz : nnz, 0);
    for (int ia = 0; ia < nnz; ia++) {
        const int j = A.indcol[ia];
        ptrowend1[ia] = seen[j] ? A.ptrow[j] : A.ptrow[j + 1];
        seen[j] = 1;
    }
}

// ---- deeper schedules of mpk/SpMVmulti0.cpp (:106, :157).  Depth d keeps one seen-set over the whole traversal:
// walking rows in order and, under each entry, the part of the neighbour row the shallower schedule left open,
// a column met for the first time at depth d opens its whole row (value = end of that row), any later meeting
// opens nothing (value = start of that row).  SpM3V / SpM4V below ignore them; they are filled so that a driver
// which builds or prints them before the product behaves as it did.  Memory is O(nnz * row^(d-1)), as there.
namespace {
inline int first_touch_end(std::vector<char> &seen, const csrmatrix &A, int col)
{
    const int e = seen[col] ? A.ptrow[col] : A.ptrow[col + 1];
    seen[col] = 1;
    return e;
}
inline int stored_entries(const csrmatrix &A) { return A.ptrow.empty() ? 0 : A.ptrow[A.n]; }
}  // namespace

void Generate2ndlayer(std::vector<std::vector<int> > &ptrowend2, csrmatrix &A, std::vector<int> &ptrowend1)
{
    std::vector<char> seen(A.n > 0 ? A.n : 1, 0);
    const int nnz = stored_entries(A);
    ptrowend2.resize(A.nnz > nnz ? A.nnz : nnz);
    for (int ia = 0; ia < nnz; ia++) {
        const int open0 = A.ptrow[A.indcol[ia]];
        const int nopen = ptrowend1[ia] - open0;
        std::vector<int> &ends = ptrowend2[ia];
        ends.resize(nopen > 0 ? nopen : 0);
        for (int t = 0; t < nopen; t++) ends[t] = first_touch_end(seen, A, A.indcol[open0 + t]);
    }
}

void Generate3rdlayer(std::vector<std::vector<std::vector<int> > > &ptrowend3, csrmatrix &A, std::vector<int> &ptrowend1,
                      std::vector<std::vector<int> > &ptrowend2)
{
    std::vector<char> seen(A.n > 0 ? A.n : 1, 0);
    const int nnz = stored_entries(A);
    ptrowend3.resize(A.nnz > nnz ? A.nnz : nnz);
    for (int ia = 0; ia < nnz; ia++) {
        const int open0 = A.ptrow[A.indcol[ia]];
        const int nopen = ptrowend1[ia] - open0;
        std::vector<std::vector<int> > &under = ptrowend3[ia];
        under.resize(nopen > 0 ? nopen : 0);
        for (int t = 0; t < nopen; t++) {
            const int k = A.indcol[open0 + t];
            const int kopen = ptrowend2[ia][t] - A.ptrow[k];
            if (kopen <= 0) continue;   // nothing left open under this entry: its list stays as it was
            std::vector<int> &ends = under[t];
            ends.resize(kopen);
            for (int u = 0; u < kopen; u++) ends[u] = first_touch_end(seen, A, A.indcol[A.ptrow[k] + u]);
        }
    }
}

void SpM2V_CSR(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &) { spm2v(z, y, x, A, NSK_EXACT_MULADD); }
void SpM2V_CSR_OPT(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &) { spm2v(z, y, x, A, NSK_EXACT_FMA); }
void SpM2V_CSR_AVX2(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &) { spm2v(z, y, x, A, NSK_FAST); }

// ---- k = 2, 3, 4 with the names and signatures of mpk/SpMVmulti0.cpp (:44, :65, :132, :191).  The nested
// first-touch schedules are accepted and ignored: the result -- every level of A^k x -- does not depend on them.
void SpMV(double *y, double *x, csrmatrix &A) { spmv(y, x, A, NSK_EXACT_MULADD); }   // the seed file's plain product (:259)
void SpM2V0(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &) { spm2v(z, y, x, A, NSK_EXACT_MULADD); }
void SpM2V(double *z, double *y, double *x, csrmatrix &A, std::vector<int> &) { spm2v(z, y, x, A, NSK_EXACT_MULADD); }
void SpM3V(double *w, double *z, double *y, double *x, csrmatrix &A, std::vector<int> &, std::vector<std::vector<int> > &)
{
    double *levels[3] = {y, z, w};
    spmkv(3, levels, x, A, NSK_EXACT_MULADD);
}
void SpM4V(double *v, double *w, double *z, double *y, double *x, csrmatrix &A, std::vector<int> &,
           std::vector<std::vector<int> > &, std::vector<std::vector<std::vector<int> > > &)
{
    double *levels[4] = {y, z, w, v};
    spmkv(4, levels, x, A, NSK_EXACT_MULADD);
}

// ---- fused A^2 x on the block operator (mpk/SpM2V.cpp:28, :376, :475, :567, :675): two block products; the schedule
// builder keeps the reference's contract (end of block row bj the first time bj is met, its start afterwards).
void Generate1stlayer_BCSR4(std::vector<int> &ptrowendB, const bcsr4x4_matrix &A)
{
    std::vector<char> seen(A.nrows > 0 ? A.nrows : 1, 0);
    ptrowendB.assign(A.indcol.size(), 0);
    for (int bi = 0; bi < A.nrows; bi++)
        for (int m = A.ptrow[bi]; m < A.ptrow[bi + 1]; m++) {
            const int bj = A.indcol[m];
            ptrowendB[m] = seen[bj] ? A.ptrow[bj] : A.ptrow[bj + 1];
            seen[bj] = 1;
        }
}
void SpM2V_BCSR(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &) { spm2v_b(z, y, x, A, NSK_EXACT_MULADD); }
void SpM2V_BCSR_OPT(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &) { spm2v_b(z, y, x, A, NSK_EXACT_FMA); }
void SpM2V_BCSR_FMA(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &) { spm2v_b(z, y, x, A, NSK_EXACT_FMA); }
void SpM2V_BCSR_AVX2(double *z, double *y, double *x, bcsr4x4_matrix &A, std::vector<int> &) { spm2v_b(z, y, x, A, NSK_EXACT_FMA); }

extern "C" void nsk_shim_reset(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &kv : g_csr) nsk_csr_destroy(kv.second.handle);
    for (auto &kv : g_bcsr) nsk_bcsr4_destroy(kv.second.handle);
    g_csr.clear();
    g_bcsr.clear();
    if (g_ctx) nsk_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}
