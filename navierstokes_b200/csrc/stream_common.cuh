// stream_common.cuh -- pieces shared by the streaming SpMV kernel and the wavefront matrix-powers kernel.
#pragma once
#include "nsk_internal.h"
#include "ptx_helpers.cuh"

// -----------------------------------------------------------------------------------------------
// arithmetic flavours
// -----------------------------------------------------------------------------------------------
template <bool MULADD>
__device__ __forceinline__ double row_op(double a, double x, double acc)
{
    if (MULADD) return __dadd_rn(acc, __dmul_rn(a, x));  // two roundings, never contracted
    return __fma_rn(a, x, acc);                          // one rounding (vfmadd231sd on the CPU)
}

// One row of y = A x out of a shared-memory stage: nonzeros j in [p,q) in storage order.
// The gathers of up to 8 consecutive nonzeros are issued together (predicated) before the dependent
// multiply-add chain starts, so a short row (stencils: 5-7 entries) costs ONE memory round trip instead
// of one per unrolled-loop remainder iteration; the chain itself stays strictly sequential.
// NC: x is constant for the whole launch (read-only path) vs written earlier in this launch (coherent).
template <bool MULADD, bool NC>
__device__ __forceinline__ double row_chain(const double *val_s, const int *col_s, int p, int q, int vo, int co,
                                            const double *src, double acc = 0.0)
{
    for (int j0 = p; j0 < q; j0 += 8) {
        double xv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int j = j0 + u;
            if (j < q) {
                const int c = col_s[j - co];
                xv[u] = NC ? __ldg(src + c) : src[c];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (j0 + u < q) acc = row_op<MULADD>(val_s[j0 + u - vo], xv[u], acc);  // coefficient straight from smem
    }
    return acc;
}

// Rows of one pass of a consumer thread: RPT rows, the gathers of their first 8 nonzeros all in flight
// before the first dependent multiply-add (the chain of each row stays strictly sequential).
template <bool MULADD, bool NC, int RPT, int NCT>
__device__ __forceinline__ void consume_tile(const double *val_s, const int *col_s, const int *ptr_s, int vo, int co,
                                             int po, int row0, int nrows, int row_end, const double *src, double *dst,
                                             int ctid)
{
    for (int rb = 0; rb < nrows; rb += NCT * RPT) {
        int p[RPT], q[RPT];
        double xv[RPT][8];
#pragma unroll
        for (int u = 0; u < RPT; u++) {
            const int r = rb + u * NCT + ctid;
            const int row = row0 + r;
            const bool valid = r < nrows && row < row_end;
            p[u] = valid ? ptr_s[row - po] : 0;
            q[u] = valid ? ptr_s[row + 1 - po] : -1;  // q < p marks "no row": nothing gathered, nothing stored
        }
#pragma unroll
        for (int u = 0; u < RPT; u++)
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int j = p[u] + e;
                xv[u][e] = 0.0;
                if (j < q[u]) {
                    const int c = col_s[j - co];
                    xv[u][e] = NC ? __ldg(src + c) : src[c];
                }
            }
#pragma unroll
        for (int u = 0; u < RPT; u++) {
            double acc = 0.0;
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (p[u] + e < q[u]) acc = row_op<MULADD>(val_s[p[u] + e - vo], xv[u][e], acc);
            if (q[u] > p[u] + 8) acc = row_chain<MULADD, NC>(val_s, col_s, p[u] + 8, q[u], vo, co, src, acc);
            if (q[u] >= p[u]) dst[row0 + rb + u * NCT + ctid] = acc;
        }
    }
}

// -----------------------------------------------------------------------------------------------
// shared-memory stage of one tile: [coef slice | indcol slice | ptrow slice | header]
// -----------------------------------------------------------------------------------------------
template <int T_NNZ, int T_ROWS>
struct StageGeom {
    static_assert(T_NNZ % 4 == 0 && T_ROWS % 4 == 0, "tile sizes must be multiples of 4");
    static constexpr int VAL_OFF = 0;
    static constexpr int VAL_BYTES = (T_NNZ + 2) * 8;
    static constexpr int COL_OFF = VAL_OFF + VAL_BYTES;
    static constexpr int COL_BYTES = (T_NNZ + 4) * 4;
    static constexpr int PTR_OFF = COL_OFF + COL_BYTES;
    static constexpr int PTR_BYTES = (T_ROWS + 8) * 4;
    static constexpr int HDR_OFF = PTR_OFF + PTR_BYTES;
    static constexpr int BYTES = HDR_OFF + 32;  // header: row0, nrows, nz0, nz1, level, tile, dep_lo, dep_hi
    static_assert(BYTES % 16 == 0, "stage must keep 16-byte alignment");
};

