// bcsr.cu -- 4x4 block CSR product (placeholder).
#include "nsk_internal.h"

NSK_API int nsk_bcsr4_create(nsk_ctx_t ctx, int nbrows, int64_t nblocks, const int *ptrow, const int *indcol,
                             const double *coef, nsk_bcsr4_t *B)
{
    (void)nbrows; (void)nblocks; (void)ptrow; (void)indcol; (void)coef; (void)B;
    nsk_set_error(ctx, "BCSR path not built");
    return NSK_ERR_UNSUPPORTED;
}
NSK_API int nsk_bcsr4_destroy(nsk_bcsr4_t B) { (void)B; return NSK_OK; }
NSK_API int nsk_spmv_bcsr4(nsk_bcsr4_t B, const double *x, double *y, nsk_mode mode, nsk_where where)
{
    (void)B; (void)x; (void)y; (void)mode; (void)where;
    return NSK_ERR_UNSUPPORTED;
}
