"""GPU parity: vector kernels (dot / norm2 / rel_error / axpy / orthogonalize / gram) and CG."""
import numpy as np
import pytest

import navierstokes_b200 as nsk
from navierstokes_b200 import matgen
from conftest import assert_bits_equal, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 1003, 100001, 3_000_001])
def test_dot_norm_rel_error(ctx, oracle_lib, n):
    rng = np.random.default_rng(n)
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    if n == 0:
        assert ctx.dot(a, b) == 0.0
        return
    # parallel (tree) summation vs the reference's sequential sum: agree to a few ulp * sqrt(n)
    exact = float(np.sum(a.astype(np.longdouble) * b.astype(np.longdouble)))
    scale = float(np.sum(np.abs(a * b))) + 1e-300
    assert abs(ctx.dot(a, b) - exact) <= 1e-14 * scale
    # the reference's norm2 is a SEQUENTIAL sum (error grows like n*eps); ours is a tree -- compare both
    # with the exact value, and with each other at the sequential sum's accuracy
    exact_n = float(np.sqrt(np.sum(a.astype(np.longdouble) ** 2)))
    assert abs(ctx.norm2(a) - exact_n) <= 1e-14 * exact_n
    assert abs(ctx.norm2(a) - oracle_lib.norm2(a)) <= 1e-12 * exact_n
    c = a + 1e-9 * b
    r_ref = oracle_lib.rel_error(a, c)
    assert abs(ctx.rel_error(a, c) - r_ref) <= 1e-12 * r_ref
    # deterministic: same call twice gives the same bits
    assert ctx.dot(a, b) == ctx.dot(a, b)


def test_golden_vector_helpers(ctx):
    g = golden("formats")
    assert abs(ctx.norm2(g["va"]) - float(g["norm2_a"])) <= 1e-14 * float(g["norm2_a"])
    assert abs(ctx.rel_error(g["va"], g["vb"]) - float(g["rel_error_ab"])) <= 1e-12 * float(g["rel_error_ab"])
    for alpha, key in ((1e-8, "orth_ac"), (0.37, "orth_ac_alpha1")):
        y = g["vc"].copy()
        beta = ctx.orthogonalize(g["va"], y, alpha)
        # beta is a tree-summed dot: y agrees with the reference's to rounding of alpha*beta*x
        assert np.max(np.abs(y - g[key])) <= 1e-15 * max(1.0, abs(beta))


def test_axpy_bitwise(ctx):
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 4097, 1_000_003):
        x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        ref = np.array([np.float64(0)] * 0)
        expect = np.empty(n)
        import math
        for i in range(min(n, 2000)):
            expect[i] = math.fma(0.3, x[i], y[i]) if hasattr(math, "fma") else 0.3 * x[i] + y[i]
        yy = y.copy()
        ctx.axpy(0.3, x, yy)
        if hasattr(math, "fma"):
            assert_bits_equal(yy[:min(n, 2000)], expect[:min(n, 2000)])
        assert np.allclose(yy, 0.3 * x + y, rtol=0, atol=1e-15)


def test_gram(ctx):
    rng = np.random.default_rng(4)
    n = 200_003
    V = [rng.uniform(-1, 1, n) for _ in range(9)]
    G = ctx.gram(V)
    M = np.stack(V)
    assert np.allclose(G, M @ M.T, rtol=1e-12, atol=1e-9)
    assert np.array_equal(G, G.T)
    dV = [ctx.to_device(v) for v in V[:5]]
    assert np.allclose(ctx.gram(dV), (M @ M.T)[:5, :5], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (24,)), ("laplace2d_5pt", (96,))])
def test_cg_matches_oracle_iteration_count(ctx, oracle_lib, gen, args):
    """Parity unpinned (the reference has no CG): compare with the textbook restatement -- same
    iteration count within +-2 (dot association differs), true residual verified on the CPU."""
    A = getattr(matgen, gen)(*args)
    xt = matgen.vec_uniform(A.n, seed=1)
    b = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, xt)
    x_ref, it_ref, rel_ref, _ = oracle_lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-8, maxit=2000)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x, it, rel, ok = dA.cg(b, tol=1e-8, maxit=2000)
    assert ok and rel <= 1e-8
    assert abs(it - it_ref) <= 2, (it, it_ref)
    assert oracle_lib.true_relres(A.ptrow, A.indcol, A.coef, b, x) <= 2e-8
    assert np.max(np.abs(x - xt)) <= 1e-6


def test_cg_maxit_reports_not_converged(ctx, oracle_lib):
    A = matgen.laplace2d_5pt(64)
    b = matgen.vec_uniform(A.n, seed=3)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x, it, rel, ok = dA.cg(b, tol=1e-12, maxit=5)
    assert not ok and it == 5 and rel > 1e-12
    # the 5 iterations themselves follow the oracle's recurrence
    _, _, _, hist = oracle_lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-12, maxit=5)
    assert abs(rel - hist[5]) <= 1e-10 * hist[5]


def test_cg_zero_rhs(ctx):
    A = matgen.laplace2d_5pt(16)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x, it, rel, ok = dA.cg(np.zeros(A.n))
    assert ok and it == 0 and np.all(x == 0)


# ---- s-step (communication-avoiding) CG ------------------------------------------------------------------
@pytest.mark.parametrize("s", [2, 3, 4])
@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (24,)), ("laplace2d_5pt", (96,)), ("laplace3d_7pt", (40, 32, 36))])
def test_sstep_cg_matches_oracle(ctx, oracle_lib, gen, args, s):
    """Parity unpinned (no CG in the reference): s-step CG against its numpy restatement (oracle.scg) and against
    classical CG -- in exact arithmetic all three produce the same iterates, so the iteration counts agree within
    +-2 and the TRUE residual (CPU oracle SpMV) meets the tolerance."""
    A = getattr(matgen, gen)(*args)
    xt = matgen.vec_uniform(A.n, seed=1)
    b = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, xt)
    _, it_cg, _, _ = oracle_lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-8, maxit=2000)
    _, it_ref, rel_ref, ok_ref = oracle_lib.scg(A.ptrow, A.indcol, A.coef, b, s=s, tol=1e-8, maxit=2000)
    assert ok_ref
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x, it, rel, ok = dA.cg(b, tol=1e-8, maxit=2000, sstep=s)
    assert ok and rel <= 1e-8
    assert abs(it - it_ref) <= 2 and abs(it - it_cg) <= 2, (it, it_ref, it_cg)
    assert oracle_lib.true_relres(A.ptrow, A.indcol, A.coef, b, x) <= 5e-8
    assert np.max(np.abs(x - xt)) <= 1e-5


def test_sstep_cg_maxit_and_zero_rhs(ctx, oracle_lib):
    A = matgen.laplace2d_5pt(64)
    b = matgen.vec_uniform(A.n, seed=3)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x, it, rel, ok = dA.cg(b, tol=1e-12, maxit=6, sstep=4)  # 6 = one full block + a block cut after 2 iterations
    assert not ok and it == 6 and rel > 1e-12
    _, _, _, hist = oracle_lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-12, maxit=6)
    assert abs(rel - hist[6]) <= 1e-8 * hist[6]
    x, it, rel, ok = dA.cg(np.zeros(A.n), sstep=4)
    assert ok and it == 0 and np.all(x == 0)


def test_sstep_cg_device_vectors_and_strategies(ctx, oracle_lib):
    """Device-resident call; the matrix-powers strategy underneath (fused packed pipeline vs k launches) does not
    change the iterates beyond rounding of the Gram sums."""
    A = matgen.laplace3d_7pt(48)
    xt = matgen.vec_uniform(A.n, seed=2)
    b = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, xt)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    db = ctx.to_device(b)
    its = []
    for strat in (0, 1):
        ctx.set_option("mpk_kernel", strat)
        dx, it, rel, ok = dA.cg(db, tol=1e-8, maxit=2000, sstep=4)
        assert ok
        assert oracle_lib.true_relres(A.ptrow, A.indcol, A.coef, b, dx.to_host()) <= 5e-8
        its.append(it)
    ctx.set_option("mpk_kernel", 0)
    assert abs(its[0] - its[1]) <= 1


def test_orthonormalize_against_basis(ctx, oracle_lib):
    """Reference helper mpk/2SpMV.cpp:13-28 on the reference's fake Krylov basis: against the fixture made by the
    compiled reference (reduction order differs: 1e-12 relative, the fast-mode bound), host and device residency."""
    g = golden("orthobasis")
    basis = [np.ascontiguousarray(b) for b in g["basis"]]
    y = g["y"].copy()
    nrm = nsk.orthonormalize_against_basis(len(y), basis, y)
    assert oracle_lib.rel_error(g["y_out"], y) <= 1e-12
    assert abs(nrm - np.linalg.norm(g["y_out"])) <= 1e-12 * nrm
    dy = ctx.to_device(g["y"])
    nrm2 = ctx.orthonormalize_against_basis([ctx.to_device(b) for b in basis], dy)
    assert oracle_lib.rel_error(g["y_out"], dy.to_host()) <= 1e-12 and abs(nrm2 - nrm) <= 1e-12 * nrm
    # empty basis: y untouched, norm returned
    z = g["y"].copy()
    assert abs(nsk.orthonormalize_against_basis(len(z), [], z) - np.linalg.norm(g["y"])) <= 1e-12 * nrm
    assert np.array_equal(z, g["y"])
