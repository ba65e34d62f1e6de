"""navierstokes_b200 -- B200-native (sm_100a) CSR SpMV / matrix-powers / CG kernels behind the
aantoine890/navierstokes ``mpk/SpMV.h`` entry points.

The product is the C-ABI shared library ``navierstokes_b200/lib/libnsk.so`` (include/nsk.h); this
package is its Python host layer (ctypes) plus workload generators.  Importing the package does not
load the library or touch a GPU; creating a ``Context`` does, and fails loudly without one.
"""
from . import matgen  # noqa: F401
from .api import (  # noqa: F401
    DEVICE, EXACT_FMA, EXACT_MULADD, FAST, HOST, Bcsr4Matrix, Context, CsrMatrix, DeviceVector, NskError,
    BuildKrylovBasis_AVX2, Generate1stlayer_BCSR4, MatMatMult_SeqBAIJ_4_AVX2, SpM2V, SpM2V0, SpM2V_BCSR, SpM2V_BCSR_AVX2, SpM2V_BCSR_FMA, SpM2V_BCSR_OPT, SpMV_BCSR, SpMV_BCSR_AVX2, SpMV_BCSR_FMA, SpMV_BCSR_OPT, bcsr4x4_matrix,
    COO2CSR, Generate1stlayer, generate_BCSR4, read_mtx, SpM2V_CSR, SpM2V_CSR_AVX2, SpM2V_CSR_OPT, SpM3V, SpM4V, SpMV_CSR, SpMV_CSR_AVX2,
    SpMV_CSR_FMA, SpMV_CSR_OPT, csrmatrix, default_context, flush_cache, norm2, orthogonalize, orthonormalize_against_basis, rel_error,
)

__version__ = "0.1.0"
