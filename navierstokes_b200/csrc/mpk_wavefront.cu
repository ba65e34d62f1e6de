// mpk_wavefront.cu -- L2-resident wavefront matrix-powers kernel (placeholder until measured).
#include "nsk_internal.h"

bool nsk_mpk_wavefront_applicable(nsk_csr_t A, int k)
{
    (void)A; (void)k;
    return false;
}

int nsk_mpk_wavefront(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                      const int *level_rows)
{
    (void)k; (void)d_x; (void)d_levels; (void)mode; (void)level_rows;
    nsk_set_error(A->ctx, "wavefront matrix-powers kernel not built");
    return NSK_ERR_UNSUPPORTED;
}
