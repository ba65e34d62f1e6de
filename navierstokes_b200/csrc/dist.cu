// dist.cu -- row-slab distributed operator: halo exchange + redundant ghost levels (placeholder).
#include "nsk_internal.h"

void nsk_dist_free(nsk_csr_t A) { (void)A; }

int nsk_dist_mpk(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode)
{
    (void)k; (void)d_x; (void)d_levels; (void)mode;
    nsk_set_error(A->ctx, "distributed operator not built");
    return NSK_ERR_UNSUPPORTED;
}

int nsk_halo_exchange_dev(nsk_csr_t A, double *xlocal, int depth)
{
    (void)xlocal; (void)depth;
    nsk_set_error(A->ctx, "distributed operator not built");
    return NSK_ERR_UNSUPPORTED;
}

NSK_API int nsk_csr_create_dist(nsk_ctx_t ctx, int n_owned, int n_rows_local, int n_cols_local, int64_t nnz,
                                const int *ptrow, const int *indcol, const double *coef, int halo_depth,
                                const int *level_rows, int n_peers, const int *peer_rank, const int *send_off,
                                const int *send_idx, const int *recv_off, const int *recv_idx, nsk_csr_t *A)
{
    (void)n_owned; (void)n_rows_local; (void)n_cols_local; (void)nnz; (void)ptrow; (void)indcol; (void)coef;
    (void)halo_depth; (void)level_rows; (void)n_peers; (void)peer_rank; (void)send_off; (void)send_idx;
    (void)recv_off; (void)recv_idx; (void)A;
    nsk_set_error(ctx, "distributed operator not built");
    return NSK_ERR_UNSUPPORTED;
}

NSK_API int nsk_halo_exchange(nsk_csr_t A, double *xlocal, int depth)
{
    if (!A) return NSK_ERR_INVALID;
    return nsk_halo_exchange_dev(A, xlocal, depth);
}

NSK_API int64_t nsk_plan_new_columns(int nrows, const int *ptrow, const int *indcol_global, int own_begin,
                                     int own_end, const int *known_sorted, int64_t n_known, int *out)
{
    (void)nrows; (void)ptrow; (void)indcol_global; (void)own_begin; (void)own_end; (void)known_sorted;
    (void)n_known; (void)out;
    return -1;
}
