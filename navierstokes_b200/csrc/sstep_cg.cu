// sstep_cg.cu -- s-step (communication-avoiding) CG (placeholder).
#include "nsk_internal.h"

int nsk_scg_device(nsk_csr_t A, const double *d_b, double *d_x, double tol, int maxit, int s, int *iters,
                   double *relres)
{
    (void)d_b; (void)d_x; (void)tol; (void)maxit; (void)s; (void)iters; (void)relres;
    nsk_set_error(A->ctx, "s-step CG not built");
    return NSK_ERR_UNSUPPORTED;
}
