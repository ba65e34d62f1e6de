"""GPU parity: nsk_spmv / nsk_mpk through the C ABI against the oracle and the golden fixtures."""
import numpy as np
import pytest

import navierstokes_b200 as nsk
from navierstokes_b200 import matgen
from conftest import CSR_CASES, DEEP_CASES, VECS, assert_bits_equal, golden

pytestmark = pytest.mark.gpu

FAST_TOL = 1e-12  # BASELINE.json north star: fast mode within 1e-12 relative (reference rel_error)


def all_kernel_configs(ctx):
    """(spmv_kernel, stream_variant) pairs: the simple kernels and every streaming geometry."""
    yield 1, 0
    yield 4, 0  # sliced-ELL tiles (refused for very ragged rows; then the CSR kernels run)
    for v in range(1, 9):
        yield 2, v
    for v in range(1, 15):
        yield 3, v  # packed format (applies to operators whose tiles read x in a few runs; else the CSR kernels run)


def select_kernel(ctx, kern, var):
    ctx.set_option("spmv_kernel", kern)
    ctx.set_option("packed_variant" if kern == 3 else "stream_variant", var)


@pytest.fixture()
def reset_options(ctx):
    yield
    for name in ("spmv_kernel", "stream_variant", "spmv_ctas_per_sm", "mpk_kernel", "wave_l2_pct", "packed_variant"):
        ctx.set_option(name, 0)
    ctx.set_option("wave_slack_pct", -1)
    ctx.set_option("pipe_interleave", 1)
    ctx.set_option("pipe_bp_global", -1)
    ctx.set_option("pipe_w0_pct", 0)


@pytest.mark.parametrize("case", CSR_CASES)
def test_spmv_exact_fma_golden(ctx, case, reset_options):
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for kern, var in all_kernel_configs(ctx):
        select_kernel(ctx, kern, var)
        for v in VECS:
            y = A.spmv(g[f"x_{v}"], mode=nsk.EXACT_FMA)
            assert_bits_equal(y, g[f"spmv_fma_{v}"], f"{case}/{v} kernel={kern} variant={var}")


@pytest.mark.parametrize("case", CSR_CASES)
def test_spmv_exact_muladd_and_fast(ctx, oracle_lib, case, reset_options):
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for kern, var in [(1, 0), (2, 1), (2, 2)]:
        select_kernel(ctx, kern, var)
        for v in VECS:
            x = g[f"x_{v}"]
            ym = A.spmv(x, mode=nsk.EXACT_MULADD)
            assert_bits_equal(ym, oracle_lib.spmv_muladd(g["ptrow"], g["indcol"], g["coef"], x), f"{case}/{v} muladd")
            yf = A.spmv(x, mode=nsk.FAST)
            ref = g[f"spmv_fma_{v}"]
            if np.any(ref):
                assert oracle_lib.rel_error(ref, yf) <= FAST_TOL


def test_reference_named_entry_points(ctx, oracle_lib):
    """Reads like mpk/2SpMV.cpp:127-293: every variant against variant 0 with rel_error."""
    g = golden("fem_baij4_m3")
    A = nsk.csrmatrix(n=len(g["ptrow"]) - 1, nnz=len(g["indcol"]), ptrow=g["ptrow"], indcol=g["indcol"],
                      coef=g["coef"])
    x = np.ones(A.n)
    base = np.zeros(A.n)
    nsk.SpMV_CSR(base, x, A)
    for fn in (nsk.SpMV_CSR_OPT, nsk.SpMV_CSR_FMA, nsk.SpMV_CSR_AVX2):
        y = np.zeros(A.n)
        fn(y, x, A)
        assert nsk.rel_error(base, y) <= 1e-12
    y = np.zeros(A.n)
    nsk.SpMV_CSR_FMA(y, g["x_uni"], A)
    assert_bits_equal(y, g["spmv_fma_uni"])
    z, y2 = np.zeros(A.n), np.zeros(A.n)
    nsk.SpM2V_CSR_OPT(z, y2, g["x_uni"], A, None)
    lv = oracle_lib.mpk(g["ptrow"], g["indcol"], g["coef"], 2, g["x_uni"])
    assert_bits_equal(y2, lv[0])
    assert_bits_equal(z, lv[1])


@pytest.mark.parametrize("n,mean", [(1, 3.0), (7, 2.0), (1000, 0.5), (5000, 5.0), (3000, 40.0), (2000, 130.0)])
def test_spmv_ragged_edge_cases(ctx, oracle_lib, n, mean, reset_options):
    """Empty rows, rows longer than a warp, tiny and single-row operators."""
    A = matgen.random_csr(n, mean, seed=n, empty_rows=True)
    x = matgen.vec_uniform(n, seed=2)
    ref = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    for kern, var in all_kernel_configs(ctx):
        select_kernel(ctx, kern, var)
        assert_bits_equal(dA.spmv(x, mode=nsk.EXACT_FMA), ref, f"n={n} kernel={kern} variant={var}")
        yf = dA.spmv(x, mode=nsk.FAST)
        if np.any(ref):
            assert oracle_lib.rel_error(ref, yf) <= FAST_TOL


def test_spmv_empty_operator(ctx):
    A = nsk.CsrMatrix(ctx, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert A.spmv(np.zeros(0)).size == 0
    B = nsk.CsrMatrix(ctx, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert np.array_equal(B.spmv(np.ones(5)), np.zeros(5))


def test_spmv_row_longer_than_a_stage(ctx, oracle_lib, reset_options):
    """One dense row of 10000 entries among short ones: exceeds every tile geometry (max 4096)."""
    n = 12000
    rng = np.random.default_rng(5)
    lens = np.full(n, 3)
    lens[17] = 10000
    lens[n - 1] = 5000
    ptrow = np.zeros(n + 1, np.int32)
    ptrow[1:] = np.cumsum(lens)
    indcol = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.int32)
    coef = rng.uniform(-1, 1, ptrow[-1])
    x = rng.uniform(-1, 1, n)
    ref = oracle_lib.spmv(ptrow, indcol, coef, x)
    dA = nsk.CsrMatrix(ctx, ptrow, indcol, coef)
    for var in (1, 2, 5):
        ctx.set_option("stream_variant", var)
        assert_bits_equal(dA.spmv(x, mode=nsk.EXACT_FMA), ref, f"variant {var}")
        assert oracle_lib.rel_error(ref, dA.spmv(x, mode=nsk.FAST)) <= FAST_TOL


def test_invalid_arguments_are_errors(ctx):
    with pytest.raises(nsk.NskError):
        nsk.CsrMatrix(ctx, np.array([0, 2, 1], np.int32), np.array([0, 1], np.int32), np.ones(2))  # decreasing
    with pytest.raises(nsk.NskError):
        nsk.CsrMatrix(ctx, np.array([0, 1], np.int32), np.array([5], np.int32), np.ones(1))  # column out of range


@pytest.mark.parametrize("case", CSR_CASES)
@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_mpk_equals_k_products_bitwise(ctx, oracle_lib, case, k):
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for v in VECS:
        lv = A.mpk(k, g[f"x_{v}"], mode=nsk.EXACT_FMA)
        assert_bits_equal(lv, oracle_lib.mpk(g["ptrow"], g["indcol"], g["coef"], k, g[f"x_{v}"]), f"{case}/{v}/k={k}")


@pytest.mark.parametrize("case", DEEP_CASES)
def test_mpk_against_reference_spmkv_golden(ctx, oracle_lib, case):
    """SpM2V0 / SpM3V / SpM4V outputs of the reference (x87 / mixed contraction): 1e-12 bound."""
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for depth in (2, 3, 4):
        for v in VECS:
            lv = A.mpk(depth, g[f"x_{v}"], mode=nsk.EXACT_FMA)
            ref = g[f"multi0_k{depth}_{v}"]
            for l in range(depth):
                assert oracle_lib.rel_error(ref[l], lv[l]) <= 1e-12
    # multiply-add flavour reproduces the no-fma reference bit for bit on the stencil case
    if case == "lap3d_7pt_6":
        lv = A.mpk(4, g["x_uni"], mode=nsk.EXACT_MULADD)
        assert_bits_equal(lv, g["multi0_k4_uni"])


def test_spm2v_opt_golden_mixed_flavour(ctx):
    """SpM2V_CSR_OPT as the reference compiles here: level 1 multiply-add, level 2 fma."""
    g = golden("lap3d_7pt_12x10x9")
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for v in VECS:
        y = A.spmv(g[f"x_{v}"], mode=nsk.EXACT_MULADD)
        z = A.spmv(y, mode=nsk.EXACT_FMA)
        assert_bits_equal(y, g[f"spm2v_opt_y_{v}"])
        assert_bits_equal(z, g[f"spm2v_opt_z_{v}"])


def test_device_resident_path_matches_host_path(ctx):
    A = matgen.laplace3d_7pt(20)
    x = matgen.vec_uniform(A.n)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    dx = ctx.to_device(x)
    lv_d = dA.mpk(4, dx)
    lv_h = dA.mpk(4, x)
    for l in range(4):
        assert_bits_equal(lv_d[l].to_host(), lv_h[l])
    assert ctx.launch_count > 0


@pytest.mark.parametrize("shape", ["2d", "3d"])
def test_medium_size_properties(ctx, oracle_lib, shape, reset_options):
    """Larger than any fixture (seconds on the oracle): bit-exact against the oracle, plus linearity
    A(ax+by) ~= aAx + bAy and the constant-vector identity of the stencil (row sums)."""
    A = matgen.laplace2d_5pt(700) if shape == "2d" else matgen.laplace3d_7pt(80)
    x, w = matgen.vec_uniform(A.n, 1), matgen.vec_sin(A.n)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    for var in (1, 2, 3, 4, 5, 6, 7, 8):
        ctx.set_option("stream_variant", var)
        y = dA.spmv(x)
        assert_bits_equal(y, oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"variant {var}")
    yw = dA.spmv(w)
    comb = dA.spmv(2.0 * x - 3.0 * w)
    assert oracle_lib.rel_error(2.0 * y - 3.0 * yw, comb) <= 1e-13
    ones = dA.spmv(np.ones(A.n))
    rowsum = np.add.reduceat(A.coef, A.ptrow[:-1])
    assert np.array_equal(ones, rowsum)


# ---- wavefront (single-launch, L2-resident) matrix powers ------------------------------------------
# ---- level-pipelined (single-launch, CTAs specialised by level) matrix powers -------------------------
# ---- packed format: SpMV and matrix powers with x runs staged in shared memory ------------------------
PACKED_OPS = [("laplace3d_7pt", (40,)), ("laplace2d_5pt", (300,)), ("laplace3d_7pt", (64, 64, 20)),
              ("laplace2d_5pt", (1000, 37)), ("laplace3d_7pt", (34, 10, 50)), ("laplace3d_7pt", (33, 7, 5)),
              ("laplace2d_5pt", (255, 3))]


@pytest.mark.parametrize("variant", list(range(1, 15)))
@pytest.mark.parametrize("gen,args", PACKED_OPS)
def test_packed_spmv_and_mpk_bitwise(ctx, oracle_lib, gen, args, variant, reset_options):
    A = getattr(matgen, gen)(*args)
    x = matgen.vec_uniform(A.n, seed=11)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("packed_variant", variant)
    ctx.set_option("spmv_kernel", 3)
    assert dA.packed_bytes > 0, "the packed format must apply to a stencil operator"
    assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"{gen}{args} spmv")
    assert_bits_equal(dA.spmv(x, mode=nsk.EXACT_MULADD), oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, x))
    ctx.set_option("mpk_kernel", 4)
    for k in (2, 4, 7):
        before = ctx.launch_count
        lv = dA.mpk(k, x, mode=nsk.EXACT_FMA)
        assert ctx.launch_count - before == 1, "packed level pipeline did not apply"
        assert_bits_equal(lv, oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x), f"{gen}{args} k={k} variant={variant}")


@pytest.mark.parametrize("interleave,bp,w0", [(1, 1, 0), (0, 1, 0), (1, 0, 0), (0, 0, 250), (1, 1, 40)])
@pytest.mark.parametrize("lead_pct", [-1, 0, 25, 100, 400])
def test_packed_mpk_lead_and_placement(ctx, oracle_lib, interleave, bp, w0, lead_pct, reset_options):
    """Window from tight (lead = reach + one group) to loose, both back-pressure modes, both level placements,
    uneven teams: same bits."""
    A = matgen.laplace3d_7pt(48, 40, 36)
    x = matgen.vec_uniform(A.n, seed=3)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("mpk_kernel", 4)
    ctx.set_option("pipe_interleave", interleave)
    ctx.set_option("pipe_bp_global", bp)
    ctx.set_option("pipe_w0_pct", w0)
    ctx.set_option("wave_slack_pct", lead_pct)
    ctx.set_option("wave_l2_pct", 1000)  # never refuse: this operator is tiny, the window bound is not the subject
    for k in (2, 5, 16):
        before = ctx.launch_count
        lv = dA.mpk(k, x)
        assert ctx.launch_count - before == 1
        assert_bits_equal(lv, oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x), f"k={k}")


def test_packed_long_rows_fem_operator(ctx, oracle_lib, reset_options):
    """FEM-like BAIJ-4 operator (58 nonzeros per row): packs with the long-row stage geometry chosen automatically;
    SpMV and fused powers bit-exact in both exact flavours."""
    A = matgen.fem_baij4(7)
    x = matgen.vec_uniform(A.n, seed=13)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    assert dA.packed_bytes > 0
    assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x))
    assert ctx.query("last_spmv_kernel") == 3  # explicit-column tiles: the packed long-row geometry is the default
    assert_bits_equal(dA.spmv(x, mode=nsk.EXACT_MULADD), oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, x))
    ctx.set_option("mpk_kernel", 4)
    for k in (2, 3):
        assert_bits_equal(dA.mpk(k, x), oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x), f"k={k}")
        assert ctx.query("last_mpk_strategy") == 4
    for variant in (11, 12, 13, 14, 8):
        ctx.set_option("packed_variant", variant)
        assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"variant {variant}")
        assert_bits_equal(dA.mpk(2, x), oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 2, x), f"variant {variant} k=2")


def test_packed_repeated_calls_are_stable(ctx, reset_options):
    A = matgen.laplace3d_7pt(56)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    dx = ctx.to_device(matgen.vec_uniform(A.n, 4))
    lv = [ctx.empty(A.n) for _ in range(4)]
    ctx.set_option("mpk_kernel", 4)
    dA.mpk(4, dx, lv)
    first = [l.to_host() for l in lv]
    for _ in range(20):
        dA.mpk(4, dx, lv)
    for l in range(4):
        assert_bits_equal(lv[l].to_host(), first[l])


def test_packed_banded_ragged_and_refusals(ctx, oracle_lib, reset_options):
    """A banded operator with mildly ragged / empty rows whose tiles need few runs packs; a Poisson-ragged banded
    operator, a random operator and one with a random operator do not -- the CSR kernels run instead, same bits; an odd-sized operator packs too."""
    ctx.set_option("spmv_kernel", 3)
    ctx.set_option("mpk_kernel", 4)
    R = matgen.random_banded_csr(20000, 150, 5.0, seed=4, empty_rows=True)
    xr = matgen.vec_uniform(R.n, seed=8)
    dR = nsk.CsrMatrix(ctx, R.ptrow, R.indcol, R.coef)
    assert dR.packed_bytes == 0  # slot-major tiles would pad Poisson row lengths by more than a third
    assert_bits_equal(dR.mpk(5, xr), oracle_lib.mpk(R.ptrow, R.indcol, R.coef, 5, xr))
    A = matgen.random_banded_csr(20000, 150, 5.0, seed=4, empty_rows=True, len_range=(4, 6))
    x = matgen.vec_uniform(A.n, seed=8)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    assert dA.packed_bytes > 0
    assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x))
    assert_bits_equal(dA.mpk(5, x), oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 5, x))
    B = matgen.random_csr(4000, 5.0, seed=1, empty_rows=True)
    xb = matgen.vec_uniform(B.n)
    dB = nsk.CsrMatrix(ctx, B.ptrow, B.indcol, B.coef)
    assert dB.packed_bytes == 0
    assert_bits_equal(dB.spmv(xb), oracle_lib.spmv(B.ptrow, B.indcol, B.coef, xb))
    assert_bits_equal(dB.mpk(3, xb), oracle_lib.mpk(B.ptrow, B.indcol, B.coef, 3, xb))
    C = matgen.laplace3d_7pt(11, 9, 7)  # 693 rows: odd -- the last element of every vector is copied by hand
    xc = matgen.vec_uniform(C.n)
    dC = nsk.CsrMatrix(ctx, C.ptrow, C.indcol, C.coef)
    assert dC.packed_bytes > 0
    assert_bits_equal(dC.spmv(xc), oracle_lib.spmv(C.ptrow, C.indcol, C.coef, xc))
    assert_bits_equal(dC.mpk(3, xc), oracle_lib.mpk(C.ptrow, C.indcol, C.coef, 3, xc))
    assert ctx.query("last_mpk_strategy") == 4


# ---- distributed operator, degenerate world of one rank (the N > 1 path is covered by the gloo tests
# ---- on CPU and by tools/dist_check.py under torchrun) ----------------------------------------------
class _OneRank:
    def get_rank(self):
        return 0

    def get_world_size(self):
        return 1

    def all_gather_object(self, out, obj):
        out[0] = obj

    def broadcast_object_list(self, objs, src=0):
        return None


def test_distributed_operator_world1(ctx, oracle_lib, reset_options):
    from navierstokes_b200 import distributed as nd
    nx, ny, nz, K = 20, 18, 16, 4
    A = matgen.laplace3d_7pt(nx, ny, nz)
    x = matgen.vec_uniform(A.n, seed=5)
    op = nd.DistStencil3D(ctx, _OneRank(), nx, ny, nz, halo_depth=K)
    assert op.n_owned == A.n and op.n_cols_local == A.n
    dx = op.new_vector()
    op.set_owned(dx, x)
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, K, x)
    for strat in (1, 2, 3, 4):
        ctx.set_option("mpk_kernel", strat)
        lv = [op.new_vector() for _ in range(K)]
        op.mpk(K, dx, lv)
        for l in range(K):
            assert_bits_equal(op.get_owned(lv[l]), ref[l], f"strategy {strat} level {l}")
    b = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x)
    xs, it, rel, ok = op.cg(b, tol=1e-9, maxit=500)
    assert ok and np.max(np.abs(xs - x)) < 1e-6
    xs4, it4, rel4, ok4 = op.cg(b, tol=1e-9, maxit=500, sstep=4)
    assert ok4 and abs(it4 - it) <= 2 and np.max(np.abs(xs4 - x)) < 1e-6


# ---- 4x4 block CSR (the reference's fastest format for its FEM matrices) -------------
def test_bcsr4_golden_bitwise(ctx):
    """SpMV_BCSR_{OPT,FMA,AVX2} on the reference's own generate_BCSR4 output (fixture made by oracle/_ref): bit
    for bit; the x87 variant within its own reproducibility bound; fused A^2 x = two products."""
    g = golden("formats")
    B = nsk.bcsr4x4_matrix(nrows=len(g["bcsr_ptrow"]) - 1, nblocks=len(g["bcsr_indcol"]), ptrow=g["bcsr_ptrow"],
                           indcol=g["bcsr_indcol"], coef=g["bcsr_coef"])
    x = g["x"]
    n = 4 * B.nrows
    for fn, key in ((nsk.SpMV_BCSR_OPT, "opt"), (nsk.SpMV_BCSR_FMA, "fma"), (nsk.SpMV_BCSR_AVX2, "avx2")):
        y = np.zeros(n)
        fn(y, x, B)
        assert_bits_equal(y, g[f"bcsr_spmv_{key}"], key)
    y = np.zeros(n)
    nsk.SpMV_BCSR(y, x, B)
    assert nsk.rel_error(g["bcsr_spmv_x87"], y) <= 1e-15
    for fn, key in ((nsk.SpM2V_BCSR_OPT, "opt"), (nsk.SpM2V_BCSR_FMA, "opt"), (nsk.SpM2V_BCSR_AVX2, "avx2"), (nsk.SpM2V_BCSR, "x87")):
        y, z = np.zeros(n), np.zeros(n)
        fn(z, y, x, B)
        assert nsk.rel_error(g[f"bcsr_spm2v_{key}_y"], y) <= 1e-15 and nsk.rel_error(g[f"bcsr_spm2v_{key}_z"], z) <= 1e-14


def test_bcsr4_fem_operator_matches_oracle_and_csr(ctx, oracle_lib):
    """FEM-like BAIJ-4 operator: blocked on the host like generate_BCSR4, product bit-identical to the oracle's
    SpMV_BCSR_FMA restatement, and equal to the CSR product up to the explicit zeros' rounding (<= 1e-13)."""
    A = matgen.fem_baij4(6)
    Bh = matgen.csr_to_bcsr4(A)
    rows = np.repeat(np.arange(A.n, dtype=np.int32), np.diff(A.ptrow))
    bp, bc, bv = oracle_lib.generate_bcsr4(A.n, rows, A.indcol, A.coef)
    assert np.array_equal(bp, Bh.ptrow) and np.array_equal(bc, Bh.indcol)
    assert_bits_equal(bv, Bh.coef)
    x = matgen.vec_uniform(A.n, seed=6)
    dB = nsk.Bcsr4Matrix(ctx, Bh.ptrow, Bh.indcol, Bh.coef)
    y = dB.spmv(x)
    assert_bits_equal(y, oracle_lib.spmv_bcsr4(Bh.ptrow, Bh.indcol, Bh.coef, x))
    dx = ctx.to_device(x)
    assert_bits_equal(dB.spmv(dx).to_host(), y)
    y_csr = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef).spmv(x)
    assert oracle_lib.rel_error(y_csr, y) <= 1e-13
    # every load-batching instance of the block kernel (block rows of 27, 18, 12, 8 blocks: batches + remainders), both flavours
    try:
        ctx.set_option("bcsr_batch", 1)
        want_ma = dB.spmv(dx, mode=nsk.EXACT_MULADD).to_host()  # the one-block-at-a-time instance (the round-1 kernel)
        assert oracle_lib.rel_error(y, want_ma) <= 1e-14
        for batch in (1, 2, 4):
            ctx.set_option("bcsr_batch", batch)
            assert_bits_equal(dB.spmv(dx).to_host(), y, f"batch {batch}")
            assert_bits_equal(dB.spmv(dx, mode=nsk.EXACT_MULADD).to_host(), want_ma, f"batch {batch} muladd")
    finally:
        ctx.set_option("bcsr_batch", 0)


# ---- BASELINE.json full sizes: size-independent properties (the CPU oracle would need minutes here; bench.py
# ---- additionally compares the 256^3 k=4 run bit for bit with the compiled reference on every round) -------------
@pytest.mark.parametrize("cfg", ["c2_2d_4096", "c3_3d_256"])
def test_full_size_properties(ctx, oracle_lib, cfg, reset_options):
    A = matgen.laplace2d_5pt(4096) if cfg == "c2_2d_4096" else matgen.laplace3d_7pt(256)
    n = A.n
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    assert dA.packed_bytes > 0
    x = matgen.vec_uniform(n, seed=31)
    w = matgen.vec_sin(n)
    dx, dw = ctx.to_device(x), ctx.to_device(w)
    # 1. every kernel family produces the same bits (thread-per-row from global = the simplest possible restatement)
    outs = {}
    for kern in (1, 2, 3, 4):
        ctx.set_option("spmv_kernel", kern)
        y = ctx.empty(n)
        dA.spmv(dx, y)
        assert ctx.query("last_spmv_kernel") == kern
        outs[kern] = y.to_host()
    ctx.set_option("spmv_kernel", 0)
    assert_bits_equal(outs[2], outs[1], "stream vs scalar")
    assert_bits_equal(outs[3], outs[1], "packed vs scalar")
    assert_bits_equal(outs[4], outs[1], "sliced-ELL vs scalar")
    # 1b. against the ORACLE at full size: 16 random slabs of 65536 rows, bit for bit (SpMV_CSR_FMA, mpk/SpMV.cpp:41-56)
    rng = np.random.default_rng(77)
    for r0 in rng.integers(0, n - 65536, 16):
        r0 = int(r0)
        p0, p1 = int(A.ptrow[r0]), int(A.ptrow[r0 + 65536])
        ref = oracle_lib.spmv((A.ptrow[r0:r0 + 65537] - p0).astype(np.int32), A.indcol[p0:p1], A.coef[p0:p1], x)
        assert_bits_equal(outs[1][r0:r0 + 65536], ref, f"{cfg} rows {r0}..{r0 + 65536} vs oracle")
    # 2. a sample of rows against the definition, evaluated with the same fma chain on the host
    rows = np.random.default_rng(5).integers(0, n, 2000)
    for r in rows:  # numpy has no fma: a few ulp of the row's magnitude instead of bits
        ref = float(np.dot(A.coef[A.ptrow[r]:A.ptrow[r + 1]], x[A.indcol[A.ptrow[r]:A.ptrow[r + 1]]]))
        assert abs(outs[1][r] - ref) <= 1e-14 * max(1.0, abs(ref)) * 8
    # 3. constant vector: interior rows of the Laplacian sum to zero exactly, boundary rows to small integers
    ones = dA.spmv(ctx.to_device(np.ones(n))).to_host()
    assert np.array_equal(ones, np.add.reduceat(A.coef, A.ptrow[:-1]))
    # 4. fused powers == k launches == repeated products, bit for bit, all four levels
    k = 4
    fused = [ctx.empty(n) for _ in range(k)]
    ctx.set_option("mpk_kernel", 0)
    dA.mpk(k, dx, fused)
    assert ctx.query("last_mpk_strategy") == 5  # stencil: the sliced-ELL level pipeline is the default
    ctx.set_option("mpk_kernel", 4)
    packed4 = [ctx.empty(n) for _ in range(k)]
    dA.mpk(k, dx, packed4)
    assert ctx.query("last_mpk_strategy") == 4
    for l in range(k):
        assert_bits_equal(packed4[l].to_host(), fused[l].to_host(), f"packed vs sliced-ELL level {l}")
    del packed4
    ctx.set_option("mpk_kernel", 1)
    ctx.set_option("spmv_kernel", 2)
    lev = [ctx.empty(n) for _ in range(k)]
    dA.mpk(k, dx, lev)
    for l in range(k):
        assert_bits_equal(fused[l].to_host(), lev[l].to_host(), f"level {l}")
    ctx.set_option("mpk_kernel", 0)
    ctx.set_option("spmv_kernel", 0)
    # 5. linearity of the fused kernel: A^4 (2x - 3w) == 2 A^4 x - 3 A^4 w to rounding
    comb = ctx.to_device(2.0 * x - 3.0 * w)
    lc = [ctx.empty(n) for _ in range(k)]
    lw = [ctx.empty(n) for _ in range(k)]
    dA.mpk(k, comb, lc)
    dA.mpk(k, dw, lw)
    lhs = lc[k - 1].to_host()
    rhs = 2.0 * fused[k - 1].to_host() - 3.0 * lw[k - 1].to_host()
    assert nsk.rel_error(rhs, lhs) <= 1e-13
    # 6. symmetry: <A x, w> == <x, A w>
    yx, yw = outs[1], dA.spmv(dw).to_host()
    assert abs(np.dot(yx, w) - np.dot(x, yw)) <= 1e-9 * abs(np.dot(yx, w))


@pytest.mark.parametrize("cfg", ["c1_fem_530k", "c4_tet_531k"])
def test_c1_c4_half_million_rows_vs_oracle(ctx, oracle_lib, cfg, reset_options):
    """BASELINE configs 1 and 4 at half a million rows against the oracle, bit for bit: the FEM-like 4-dof-per-node
    operator (58 per row) and the RCM-ordered tetrahedral P1 Laplacian (15 per row, unstructured): product in every
    kernel family, fused powers (k = 4 / k = 8) in the default strategy and as k products."""
    if cfg == "c1_fem_530k":
        A, k = matgen.fem_baij4(50), 4
    else:
        A, k = matgen.tet_p1_laplacian(80, permute_seed=2, rcm=True), 8
    assert A.n > 500000
    x = matgen.vec_uniform(A.n, seed=41)
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    dx = ctx.to_device(x)
    for kern in (1, 2, 3, 4, 0):
        ctx.set_option("spmv_kernel", kern)
        assert_bits_equal(dA.spmv(dx).to_host(), ref[0], f"{cfg} spmv kernel {kern}")
    ctx.set_option("spmv_kernel", 0)
    for strategy in (0, 1, 4, 5):
        ctx.set_option("mpk_kernel", strategy)
        lv = dA.mpk(k, dx)
        assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"{cfg} k={k} strategy {strategy}")


def test_flush_l2_keeps_results(ctx, oracle_lib):
    """nsk_flush_l2 (the reference's flush_cache, mpk/utils.cpp:146-154): scrubs the cache, touches no operand."""
    A = matgen.laplace3d_7pt(64)
    x = matgen.vec_uniform(A.n, seed=5)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    dx = ctx.to_device(x)
    y0 = dA.spmv(dx).to_host()
    before = ctx.launch_count
    ctx.flush_l2()
    ctx.sync()
    assert ctx.launch_count >= before
    nsk.flush_cache()  # the reference-named entry point
    assert_bits_equal(dA.spmv(dx).to_host(), y0)
    assert_bits_equal(y0, oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x))


def test_packed_mpk_splits_when_the_window_does_not_fit(ctx, oracle_lib, reset_options):
    """With an L2 budget too small for k levels the powers call fuses as many levels per launch as fit (here pairs or
    single products) instead of giving up on fusion altogether; same bits."""
    A = matgen.laplace3d_7pt(64, 64, 48)
    x = matgen.vec_uniform(A.n, seed=14)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 5, x)
    dx = ctx.to_device(x)
    seen = set()
    for pct in (1, 2, 3, 5, 100):
        ctx.set_option("wave_l2_pct", pct)
        lv = [ctx.empty(A.n) for _ in range(5)]
        before = ctx.launch_count
        dA.mpk(5, dx, lv)
        seen.add(ctx.launch_count - before)
        for l in range(5):
            assert_bits_equal(lv[l].to_host(), ref[l], f"budget {pct}% level {l}")
    assert 1 in seen and len(seen) >= 2, seen  # the generous budget fuses everything; a tight one needs more launches


@pytest.mark.parametrize("seed", list(range(8)))
def test_packed_fuzz_random_stencils(ctx, oracle_lib, seed, reset_options):
    """Random stencil shapes (up to 13 points within +-2 in every direction), random coefficients, truncated and
    randomly thinned rows, random (even-sized) grids: whatever packs must give the oracle's bits for SpMV and for the
    fused powers in both exact flavours; whatever does not pack still runs (CSR kernels) with the same bits."""
    rng = np.random.default_rng(1000 + seed)
    nx, ny, nz = int(rng.integers(5, 70)), int(rng.integers(3, 40)), int(rng.integers(2, 24))
    A = matgen.random_stencil3d(nx, ny, nz, seed=seed, max_points=int(rng.integers(3, 14)), drop=float(rng.uniform(0, 0.2)))
    x = matgen.vec_uniform(A.n, seed=seed + 50)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    packed = dA.packed_bytes > 0
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 4, x)
    ctx.set_option("spmv_kernel", 3)
    ctx.set_option("mpk_kernel", 4)
    assert_bits_equal(dA.spmv(x), ref[0], f"grid {nx}x{ny}x{nz} packed={packed}")
    assert ctx.query("last_spmv_kernel") == (3 if packed else 2)
    ctx.set_option("wave_l2_pct", 1000)
    dlv = dA.mpk(4, ctx.to_device(x), [ctx.empty(A.n) for _ in range(4)])  # device-resident: the fused path
    assert ctx.query("last_mpk_strategy") == (4 if packed else 1)
    assert_bits_equal(np.stack([l.to_host() for l in dlv]), ref, f"grid {nx}x{ny}x{nz} packed={packed} k=4")
    assert_bits_equal(dA.mpk(4, x), ref, "host-pointer call (levels copied out while the next ones are computed)")
    assert_bits_equal(dA.mpk(3, x, mode=nsk.EXACT_MULADD)[2],
                      oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, oracle_lib.spmv_muladd(
                          A.ptrow, A.indcol, A.coef, oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, x))))


# ---- several right-hand sides: two vectors per fused launch -------------------------------------------------------
@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (40,)), ("laplace2d_5pt", (300,)), ("laplace3d_7pt", (64, 64, 20)),
                                      ("fem_baij4", (6,)), ("random_csr", (4000, 5.0, 1))])
def test_mpk_multi_bitwise(ctx, oracle_lib, gen, args, reset_options):
    """nsk_mpk_multi: powers of 3 right-hand sides (one fused pair + one single) equal the oracle's bits per vector,
    device-resident and host-pointer calls, for operators that take the two-vector kernel (short-row stencils), the
    long-row geometry (vector by vector) and the CSR kernels (random pattern)."""
    A = getattr(matgen, gen)(*args)
    xs = [matgen.vec_uniform(A.n, seed=20 + v) for v in range(3)]
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    for k in (1, 2, 4):
        ref = [oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x) for x in xs]
        lv = dA.mpk_multi(k, [ctx.to_device(x) for x in xs])
        for v in range(3):
            for l in range(k):
                assert_bits_equal(lv[v][l].to_host(), ref[v][l], f"{gen} k={k} vector {v} level {l}")
        host = dA.mpk_multi(k, xs)
        for v in range(3):
            assert_bits_equal(host[v], ref[v], f"{gen} k={k} host vector {v}")


def test_mpk_multi_pair_is_one_launch_and_muladd(ctx, oracle_lib, reset_options):
    A = matgen.laplace3d_7pt(48)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    xs = [matgen.vec_uniform(A.n, seed=3), matgen.vec_sin(A.n)]
    dxs = [ctx.to_device(x) for x in xs]
    lv = dA.mpk_multi(4, dxs)
    before = ctx.launch_count
    dA.mpk_multi(4, dxs, lv)
    assert ctx.launch_count - before == 1  # both vectors, four levels, one launch
    lm = dA.mpk_multi(3, dxs, mode=nsk.EXACT_MULADD)
    for v in range(2):
        y = xs[v]
        for l in range(3):
            y = oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, y)
            assert_bits_equal(lm[v][l].to_host(), y, f"muladd vector {v} level {l}")


@pytest.mark.parametrize("s", [1, 3, 4, 6, 9])
def test_bcsr4_spmm_matches_the_restated_reference(ctx, oracle_lib, s):
    """Y = A X on s dense columns (MatMatMult_SeqBAIJ_4_AVX2, src/kernels/spmm_avx2.c:7-109): the product bit for bit in the
    reference's grouping (per block a four-link fma chain, then an add), and 4 * Y == the literal restatement including
    the reference's lane reduction; also through the reference-named layer and against a dense product."""
    A = matgen.fem_baij4(6)
    Bh = matgen.csr_to_bcsr4(A)
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.uniform(-1.0, 1.0, size=(A.n, s)))
    B = nsk.Bcsr4Matrix(ctx, Bh.ptrow, Bh.indcol, Bh.coef)
    Y = B.spmm(X)
    ref = oracle_lib.spmm_baij4(Bh.ptrow, Bh.indcol, Bh.coef, X)
    assert_bits_equal(Y, ref, f"spmm s={s}")
    assert_bits_equal(4.0 * Y, oracle_lib.spmm_baij4(Bh.ptrow, Bh.indcol, Bh.coef, X, literal=True), "literal (factor 4)")
    for k in range(s):  # same operator, different association than the single-chain product: close, not equal
        assert nsk.rel_error(oracle_lib.spmv_bcsr4(Bh.ptrow, Bh.indcol, Bh.coef, X[:, k].copy()), Y[:, k].copy()) <= 1e-14
    Bn = nsk.bcsr4x4_matrix(nrows=Bh.nbrows, nblocks=len(Bh.indcol), ptrow=Bh.ptrow, indcol=Bh.indcol, coef=Bh.coef)
    Y2 = np.zeros_like(Y)
    nsk.MatMatMult_SeqBAIJ_4_AVX2(Bn, X, Y2, s)
    assert_bits_equal(Y2, ref, "reference-named layer")


def test_bcsr4_krylov_basis(ctx, oracle_lib):
    """[v0, A v0, ..., A^s v0] (BuildKrylovBasis_AVX2, src/kernels/spmm_avx2.c:112-168): column j + 1 is the one-column
    product of column j, bit for bit; device-resident call equals the host call."""
    A = matgen.fem_baij4(5)
    Bh = matgen.csr_to_bcsr4(A)
    v0 = matgen.vec_uniform(A.n, seed=9)
    B = nsk.Bcsr4Matrix(ctx, Bh.ptrow, Bh.indcol, Bh.coef)
    s = 5
    V = B.krylov_basis(v0, s)
    assert_bits_equal(V[:, 0].copy(), v0)
    for j in range(s):
        nxt = oracle_lib.spmm_baij4(Bh.ptrow, Bh.indcol, Bh.coef, V[:, j:j + 1])
        assert_bits_equal(V[:, j + 1].copy(), nxt[:, 0].copy(), f"basis column {j + 1}")
    Bn = nsk.bcsr4x4_matrix(nrows=Bh.nbrows, nblocks=len(Bh.indcol), ptrow=Bh.ptrow, indcol=Bh.indcol, coef=Bh.coef)
    assert_bits_equal(nsk.BuildKrylovBasis_AVX2(Bn, v0, s), V)


@pytest.mark.parametrize("k", [2, 3, 5])
def test_bcsr4_powers_block_products_and_fused_expansion(ctx, oracle_lib, k):
    """nsk_bcsr4_mpk (SpM2V_BCSR* for k = 2, mpk/SpM2V.cpp:376-801): k block products by default, ONE fused launch of the
    level pipeline on the scalar expansion with mpk_kernel = 5 -- both bit-identical to k x SpMV_BCSR_FMA of the oracle,
    host and device calls, both exact flavours agree level by level with repeated products."""
    A = matgen.fem_baij4(7)
    Bh = matgen.csr_to_bcsr4(A)
    x = matgen.vec_uniform(A.n, seed=13)
    ref, src = [], x
    for _ in range(k):
        src = oracle_lib.spmv_bcsr4(Bh.ptrow, Bh.indcol, Bh.coef, src)
        ref.append(src)
    B = nsk.Bcsr4Matrix(ctx, Bh.ptrow, Bh.indcol, Bh.coef)
    try:
        for strat in (0, 5):
            ctx.set_option("mpk_kernel", strat)
            before = ctx.launch_count
            lv = B.mpk(k, ctx.to_device(x))
            launches = ctx.launch_count - before
            assert launches == (k if strat == 0 else 1), f"strategy {strat}: {launches} launches"
            assert ctx.query("last_mpk_strategy") == (1 if strat == 0 else 5)
            for l in range(k):
                assert_bits_equal(lv[l].to_host(), ref[l], f"strategy {strat} level {l}")
            host = B.mpk(k, x)
            for l in range(k):
                assert_bits_equal(host[l], ref[l], f"host call, strategy {strat} level {l}")
        ctx.set_option("mpk_kernel", 0)
        ma = B.mpk(k, x, mode=nsk.EXACT_MULADD)
        src = x
        for l in range(k):
            src = B.spmv(src, mode=nsk.EXACT_MULADD)
            assert_bits_equal(ma[l], src, f"muladd level {l}")
    finally:
        ctx.set_option("mpk_kernel", 0)
