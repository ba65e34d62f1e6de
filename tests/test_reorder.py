"""nsk_rcm / nsk_csr_permute (csrc/reorder.cpp, host only): a valid permutation, the permuted operator is the same
operator (P A P^T, checked through the oracle's product), and the bandwidth is in the league of scipy's RCM."""
import ctypes as C

import numpy as np
import pytest

from navierstokes_b200 import _lib, matgen


def rcm(A):
    lib = _lib.load()
    perm = np.empty(A.nrows, np.int32)
    assert lib.nsk_rcm(A.nrows, A.ptrow.ctypes.data, A.indcol.ctypes.data, perm.ctypes.data) == 0
    return perm


@pytest.mark.parametrize("gen,args", [("tet_p1_laplacian", (9, 2)), ("laplace3d_7pt", (12, 9, 7)), ("fem_baij4", (4,)),
                                      ("random_csr", (1500, 4.0, 3))])
def test_rcm_is_a_permutation_and_permute_keeps_the_operator(oracle_lib, gen, args):
    A = getattr(matgen, gen)(*args)
    perm = rcm(A)
    assert np.array_equal(np.sort(perm), np.arange(A.nrows))
    B = matgen.rcm_reorder(A)
    assert B.nnz == A.nnz and np.all(np.diff(B.ptrow) == np.diff(A.ptrow)[perm])
    for r in range(B.nrows):  # columns ascending within each row (what the reference's generate_CSR produces)
        assert np.all(np.diff(B.indcol[B.ptrow[r]:B.ptrow[r + 1]]) > 0) or gen == "random_csr"
    x = matgen.vec_uniform(A.nrows, seed=5)
    ya = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x)
    yb = oracle_lib.spmv(B.ptrow, B.indcol, B.coef, x[perm])
    assert np.allclose(yb, ya[perm], rtol=1e-13, atol=1e-13)  # same rows, entries re-sorted: equal up to rounding order


def test_rcm_bandwidth_matches_scipy_league():
    """Randomly numbered tetrahedral mesh: RCM must bring the bandwidth down by orders of magnitude, to within 1.3 x of
    scipy's reverse_cuthill_mckee (the permutations themselves differ: tie-breaking and start nodes)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    A = matgen.tet_p1_laplacian(16, permute_seed=2)
    bw0 = matgen.bandwidth(A)
    B = matgen.rcm_reorder(A)
    bw = matgen.bandwidth(B)
    S = csr_matrix((A.coef, A.indcol, A.ptrow), shape=(A.nrows, A.nrows))
    p = reverse_cuthill_mckee(S, symmetric_mode=True)
    Sp = S[p][:, p].tocoo()
    bw_scipy = int(np.max(np.abs(Sp.row - Sp.col)))
    assert bw * 8 < bw0
    assert bw <= 1.3 * bw_scipy, (bw, bw_scipy)


def test_rcm_handles_disconnected_components_and_unsymmetric_patterns():
    # two disconnected chains + an isolated node, stored with one-directional entries only
    n = 11
    rows = [[1], [2], [3], [], [5], [6], [7], [], [], [10], []]
    ptrow = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    indcol = np.array([c for r in rows for c in r], np.int32)
    A = matgen.Csr(n=n, ptrow=ptrow, indcol=indcol, coef=np.ones(len(indcol)), ncols=n)
    perm = rcm(A)
    assert np.array_equal(np.sort(perm), np.arange(n))
    B = matgen.rcm_reorder(A)
    assert matgen.bandwidth(B) == 1


def test_permute_rejects_a_non_permutation():
    lib = _lib.load()
    A = matgen.laplace2d_5pt(4, 3)
    bad = np.zeros(A.nrows, np.int32)
    p2 = np.empty(A.nrows + 1, np.int32)
    c2 = np.empty(A.nnz, np.int32)
    v2 = np.empty(A.nnz)
    assert lib.nsk_csr_permute(A.nrows, A.ptrow.ctypes.data, A.indcol.ctypes.data, A.coef.ctypes.data, bad.ctypes.data,
                               p2.ctypes.data, c2.ctypes.data, v2.ctypes.data) < 0
