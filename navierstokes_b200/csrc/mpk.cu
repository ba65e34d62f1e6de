// mpk.cu -- matrix-powers driver: levels[l] = A^(l+1) x, l = 0..k-1.
//
// Replaces the reference's first-touch fused kernels SpM2V_CSR* (mpk/SpM2V.cpp:80-332), SpM3V and
// SpM4V (mpk/SpMVmulti0.cpp:132-221) and their schedule builders Generate{1st,2nd,3rd}layer.
// The reference algorithm is a serial lazy traversal; only its RESULT is shared:
// every level equals k successive row-sequential products, which is what the exact modes
// reproduce bit for bit (per-row nonzero order is never changed, only the schedule).
//
// Strategy 1 ("levels"): k launches of the product kernel back to back on one stream; the operator is re-read from
//   HBM k times.
// Strategy 5 (default for operators made of pattern tiles -- stencils, regular bands): sell.cu, one persistent launch,
//   CTAs specialised by level and chained by completion counters with window back-pressure so that the k - 1
//   re-reads of every tile come from L2; coefficients staged by bulk copies, no per-entry column index.
// Strategy 4 (default for other operators that pack -- block-banded FEM rows): packed.cu, the same level pipeline
//   with whole tiles (coefficients, local columns, x runs) staged in shared memory.
// (Strategies 2 and 3 of round 1 -- the skewed wavefront and the level pipeline over plain CSR with global gathers --
// lost to k launches wherever they were measured and are gone.)
#include <algorithm>

#include "nsk_internal.h"

int nsk_dist_mpk(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode);  // dist.cu

int nsk_mpk_levels(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                   const int *level_rows)
{
    const double *src = d_x;
    for (int l = 0; l < k; l++) {
        nsk_spmv_args a;
        a.x = src;
        a.y = d_levels[l];
        a.row_begin = 0;
        a.row_end = level_rows ? level_rows[l] : A->n;
        a.mode = mode;
        NSK_TRY(nsk_launch_spmv(A, a));
        src = d_levels[l];
    }
    return NSK_OK;
}

// Picks the strategy: option mpk_kernel = 0 auto, 1 levels, 4 level pipeline on the packed format (packed.cu), 5 level
// pipeline on sliced-ELL tiles (sell.cu).  A fused
// kernel that does not apply to the pattern (window larger than its L2 budget, long rows) degrades to k
// launches -- still a GPU path.  level_rows: distributed slabs evaluate level l on a row prefix.
int nsk_mpk_local(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                  const int *level_rows)
{
    nsk_ctx_t ctx = A->ctx;
    int sel = (int)ctx->opt.mpk_kernel;
    const bool automatic = sel == 0;
    // default: the sliced-ELL level pipeline for operators made of pattern tiles (stencils, regular bands: staged
    // coefficients, no per-entry index), else the packed one when the operator packs, else k products
    bool explicit_auto = false;
    if (automatic) {
        sel = 4;
        if (k > 1 && nsk_sell_uniform(A)) sel = 5;
        // Unstructured operators (explicit-column tiles whose columns fit a stage: ~15-23 entries per row): fused when at
        // least min(k, 4) levels fit one launch's L2 window.  Measured on the RCM-ordered tetrahedral P1 Laplacian, k = 8
        // (profiles/r02_c4_variants.txt, r02_configs.txt): 1.03 M rows 0.260 ms fused vs 0.296 ms as 8 products (1.14x),
        // 8.1 M rows 1.70 vs 1.89 ms (1.12x, two launches of four levels), 49.8 M rows 11.6 vs 11.8 ms (three levels per
        // launch at most: no gain, so k products).  The product itself runs at the HBM copy rate there; the fused kernel
        // is bound by the latency of the uncoalesced x gathers, not by HBM.
        // Small operators stay with products too (fill and drain of the level pipeline: a 0.59 M-row slab of a 2-GPU split
        // runs k = 8 in 0.204 ms fused against 0.187 ms as products, profiles/r02_dist_c4_2gpu.txt): threshold 1 M rows,
        // option mpk_auto_explicit = rows (> 0) moves it, < 0 switches the rule off.
        else if (k > 1 && ctx->opt.mpk_auto_explicit >= 0 &&
                 A->n >= (ctx->opt.mpk_auto_explicit > 0 ? ctx->opt.mpk_auto_explicit : 1000000) && nsk_sell_explicit_staged(A)) {
            sel = 5;
            explicit_auto = true;
        }
    }
    const bool sell = sel == 5 && k > 1 && nsk_sell_applicable(A);
    if (sel == 5 && !sell) sel = 4;
    if (sell || (sel == 4 && k > 1 && nsk_packed_applicable(A))) {
        // Fuse as many levels per launch as the L2 window allows: a pattern whose reach is large (512^2-row planes)
        // may fit two levels but not four -- then A^4 x runs as two fused pairs instead of four products.
        int done = 0;
        const double *src = d_x;
        bool any_fused = false;
        while (done < k) {
            const int left = k - done;
            int took = 0;
            const int lo = (explicit_auto && done == 0) ? std::min(k, 4) : 2;
            for (int kk = left; kk >= lo && !took; kk--) {
                int s = sell ? nsk_sell_run(A, kk, src, d_levels + done, mode, level_rows ? level_rows + done : nullptr, nullptr, -1)
                             : nsk_packed_run(A, kk, src, d_levels + done, mode, level_rows ? level_rows + done : nullptr, nullptr, -1);
                if (s == NSK_OK) took = kk;
                else if (s != NSK_ERR_UNSUPPORTED) return s;
            }
            if (!took && explicit_auto && done == 0) {  // too few levels per launch to pay: k products
                ctx->last_mpk = 1;
                return nsk_mpk_levels(A, k, d_x, d_levels, mode, level_rows);
            }
            if (!took) {  // a single level (or nothing fits): one product
                nsk_spmv_args a;
                a.x = src;
                a.y = d_levels[done];
                a.row_begin = 0;
                a.row_end = level_rows ? level_rows[done] : A->n;
                a.mode = mode;
                NSK_TRY(nsk_launch_spmv(A, a));
                took = 1;
            } else {
                any_fused = true;
            }
            src = d_levels[done + took - 1];
            done += took;
        }
        ctx->last_mpk = any_fused ? (sell ? 5 : 4) : 1;
        return NSK_OK;
    }
    ctx->last_mpk = 1;
    return nsk_mpk_levels(A, k, d_x, d_levels, mode, level_rows);
}

// Two right-hand sides at once (s-step methods: powers of p and of r in the same outer step): one fused launch per
// chunk of levels reads every tile's blob once for both; anything the two-vector kernel does not cover runs the
// vectors one after the other.
int nsk_mpk_local2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                   double *const *d_levels2, nsk_mode mode, const int *level_rows)
{
    nsk_ctx_t ctx = A->ctx;
    int sel = (int)ctx->opt.mpk_kernel;
    if (sel == 0 && nsk_sell_uniform(A)) sel = 5;
    const bool sell = sel == 5 && nsk_sell_applicable(A);
    if (sell || ((sel == 0 || sel == 4 || sel == 5) && nsk_packed_applicable(A))) {
        int done = 0;
        const double *src = d_x, *src2 = d_x2;
        bool ok = true, any = false;
        while (done < k && ok) {
            const int left = k - done;
            int took = 0;
            for (int kk = left; kk >= 1 && !took; kk--) {
                int s = sell ? nsk_sell_run2(A, kk, src, d_levels + done, src2, d_levels2 + done, mode,
                                             level_rows ? level_rows + done : nullptr)
                             : nsk_packed_run2(A, kk, src, d_levels + done, src2, d_levels2 + done, mode,
                                               level_rows ? level_rows + done : nullptr);
                if (s == NSK_OK) took = kk;
                else if (s != NSK_ERR_UNSUPPORTED) return s;
            }
            if (!took) { ok = false; break; }
            any = true;
            src = d_levels[done + took - 1];
            src2 = d_levels2[done + took - 1];
            done += took;
        }
        if (ok) { ctx->last_mpk = any ? (sell ? 5 : 4) : 1; return NSK_OK; }
        // partial progress is fine: finish the remaining levels vector by vector
        if (done > 0) {
            NSK_TRY(nsk_mpk_local(A, k - done, d_levels[done - 1], d_levels + done, mode, level_rows ? level_rows + done : nullptr));
            return nsk_mpk_local(A, k - done, d_levels2[done - 1], d_levels2 + done, mode, level_rows ? level_rows + done : nullptr);
        }
    }
    NSK_TRY(nsk_mpk_local(A, k, d_x, d_levels, mode, level_rows));
    return nsk_mpk_local(A, k, d_x2, d_levels2, mode, level_rows);
}

int nsk_dist_mpk2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                  double *const *d_levels2, nsk_mode mode);  // dist.cu

int nsk_mpk_device2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                    double *const *d_levels2, nsk_mode mode)
{
    if (A->dist) return nsk_dist_mpk2(A, k, d_x, d_levels, d_x2, d_levels2, mode);
    return nsk_mpk_local2(A, k, d_x, d_levels, d_x2, d_levels2, mode, nullptr);
}

int nsk_mpk_device(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode)
{
    if (A->dist) return nsk_dist_mpk(A, k, d_x, d_levels, mode);
    return nsk_mpk_local(A, k, d_x, d_levels, mode, nullptr);
}
