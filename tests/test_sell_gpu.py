"""GPU parity of the sliced-ELL path (csrc/sell.cu): nsk_spmv / nsk_mpk / nsk_mpk_multi through the C ABI against the
oracle and the golden fixtures -- bit-exact in both exact modes, stencil (pattern tiles) and unstructured (explicit
columns) operators, every number of tiles per item."""
import numpy as np
import pytest

import navierstokes_b200 as nsk
from navierstokes_b200 import matgen
from conftest import CSR_CASES, VECS, assert_bits_equal, golden

pytestmark = pytest.mark.gpu

SELL_OPTS = ("spmv_kernel", "mpk_kernel", "sell_chunk", "sell_ctas_per_sm", "sell_max_ctas", "sell_geom", "sell_stream", "sell_rows", "sell_tma", "sell_pf_dist", "wave_l2_pct",
             "pipe_w0_pct")


@pytest.fixture()
def sell(ctx):
    ctx.set_option("spmv_kernel", 4)
    ctx.set_option("mpk_kernel", 5)
    yield ctx
    for name in SELL_OPTS:
        ctx.set_option(name, 0)
    ctx.set_option("sell_flags", -1)
    ctx.set_option("wave_slack_pct", -1)
    ctx.set_option("pipe_interleave", 1)


@pytest.mark.parametrize("case", CSR_CASES)
def test_sell_spmv_golden(sell, oracle_lib, case):
    ctx = sell
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    for chunk, stream, rows, tma in ((1, 0, 1, 0), (4, 0, 0, 3), (3, 0, 0, 6), (2, -1, 0, -1), (3, -1, 0, -1), (4, 2, 2, -1),
                                     (3, 3, 2, -1), (2, 2, 1, -1)):
        ctx.set_option("sell_chunk", chunk)
        ctx.set_option("sell_tma", tma)        # all-pattern operators: staged-coefficient kernel (n stages), -1: register kernels
        ctx.set_option("sell_rows", rows)      # streaming kernel: rows of a tile per consumer thread
        ctx.set_option("sell_stream", stream)  # 0 / n: streaming consumers on all-pattern operators; -1: an item at a time
        for v in VECS:
            x = g[f"x_{v}"]
            y = A.spmv(x, mode=nsk.EXACT_FMA)
            assert ctx.query("last_spmv_kernel") == 4, "the sliced-ELL kernel did not run"
            assert_bits_equal(y, g[f"spmv_fma_{v}"], f"{case}/{v} chunk={chunk} stream={stream} tma={tma}")
            assert_bits_equal(A.spmv(x, mode=nsk.EXACT_MULADD),
                              oracle_lib.spmv_muladd(g["ptrow"], g["indcol"], g["coef"], x), f"{case}/{v} muladd")


@pytest.mark.parametrize("case", CSR_CASES)
@pytest.mark.parametrize("k", [2, 3, 4, 8])
def test_sell_mpk_golden_equals_k_products_bitwise(sell, oracle_lib, case, k):
    ctx = sell
    g = golden(case)
    A = nsk.CsrMatrix(ctx, g["ptrow"], g["indcol"], g["coef"])
    ctx.set_option("wave_l2_pct", 1000)
    x = g["x_uni"]
    ref = oracle_lib.mpk(g["ptrow"], g["indcol"], g["coef"], k, x)
    dx = ctx.to_device(x)
    for chunk, stream, tma in ((1, 0, 0), (4, 0, 3), (2, 0, -1), (3, -1, -1), (4, 2, -1)):
        ctx.set_option("sell_chunk", chunk)
        ctx.set_option("sell_stream", stream)
        ctx.set_option("sell_tma", tma)
        lv = [ctx.zeros(A.n) for _ in range(k)]
        before = ctx.launch_count
        A.mpk(k, dx, lv)
        assert ctx.query("last_mpk_strategy") == 5 and ctx.launch_count - before == 1, "one fused sliced-ELL launch expected"
        assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"{case} k={k} chunk={chunk}")


SELL_OPS = [("laplace3d_7pt", (40,)), ("laplace2d_5pt", (300,)), ("laplace3d_7pt", (64, 64, 20)), ("laplace2d_5pt", (1000, 37)),
            ("laplace3d_7pt", (34, 10, 50)), ("laplace3d_7pt", (33, 7, 5)), ("laplace2d_5pt", (255, 3)),
            ("tet_p1_laplacian", (24, 2, True)), ("fem_baij4", (7,)), ("random_banded_csr", (30000, 700, 9.0, 3))]


@pytest.mark.parametrize("chunk,stream,rows,tma", [(0, 0, 0, 0), (1, 0, 0, 3), (4, 0, 0, 6), (2, -1, 0, -1), (4, 2, 2, -1),
                                                   (3, 3, 2, -1), (1, 2, 1, -1),
                                                   (0, 0, 0, 8), (2, 0, 0, 17)])  # explicit tiles: 8 entries per round trip / 16 with 3 CTAs
@pytest.mark.parametrize("gen,args", SELL_OPS)
def test_sell_spmv_and_mpk_bitwise(sell, oracle_lib, gen, args, chunk, stream, rows, tma):
    """Stencils (pattern tiles, no per-entry index), an RCM-ordered tetrahedral P1 Laplacian, a 4-dof-per-node FEM operator
    (58 per row) and a ragged banded matrix (explicit tiles with per-slice widths): product and fused powers, both
    exact flavours, several repetitions (the completion counters are monotone over launches)."""
    ctx = sell
    A = getattr(matgen, gen)(*args)
    x = matgen.vec_uniform(A.n, seed=11)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("sell_chunk", chunk)
    ctx.set_option("sell_stream", stream)
    ctx.set_option("sell_rows", rows)
    ctx.set_option("sell_tma", tma)
    ctx.set_option("wave_l2_pct", 1000)
    assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"{gen}{args} spmv")
    applies = ctx.query("last_spmv_kernel") == 4
    # the randomly ragged band pads its 32-row slices past the format's limit: refused, the CSR kernels run (same bits)
    assert applies or gen == "random_banded_csr"
    assert_bits_equal(dA.spmv(x, mode=nsk.EXACT_MULADD), oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, x))
    dx = ctx.to_device(x)
    for k in (2, 4, 7):
        ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
        lv = [ctx.zeros(A.n) for _ in range(k)]
        for rep in range(3):
            before = ctx.launch_count
            dA.mpk(k, dx, lv)
            if applies:
                assert ctx.launch_count - before == 1 and ctx.query("last_mpk_strategy") == 5
            assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"{gen}{args} k={k} chunk={chunk} rep={rep}")
    y = x
    lm = dA.mpk(3, x, mode=nsk.EXACT_MULADD)
    for l in range(3):
        y = oracle_lib.spmv_muladd(A.ptrow, A.indcol, A.coef, y)
        assert_bits_equal(lm[l], y, f"{gen} muladd level {l}")


@pytest.mark.parametrize("interleave,w0,cps", [(1, 0, 0), (0, 0, 0), (1, 250, 0), (1, 40, 2), (0, 0, 1)])
@pytest.mark.parametrize("lead_pct", [-1, 0, 25, 400])
def test_sell_mpk_window_and_placement(sell, oracle_lib, interleave, w0, cps, lead_pct):
    """Tightest to loosest window, both CTA placements, uneven teams, fewer CTAs per SM: same bits, no deadlock."""
    ctx = sell
    A = matgen.laplace3d_7pt(64, 64, 24)
    x = matgen.vec_uniform(A.n, seed=5)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("pipe_interleave", interleave)
    ctx.set_option("pipe_w0_pct", w0)
    ctx.set_option("sell_ctas_per_sm", cps)
    ctx.set_option("wave_slack_pct", lead_pct)
    ctx.set_option("wave_l2_pct", 1000)
    dx = ctx.to_device(x)
    for k in (2, 5):
        ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
        for chunk, stream, tma in ((1, 0, 0), (3, 0, 6), (2, -1, -1), (4, 3, -1)):
            ctx.set_option("sell_chunk", chunk)
            ctx.set_option("sell_stream", stream)
            ctx.set_option("sell_tma", tma)
            lv = dA.mpk(k, dx)
            assert ctx.query("last_mpk_strategy") == 5
            assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"k={k} chunk={chunk}")


def test_sell_flags_prefetch_and_hints(sell, oracle_lib):
    ctx = sell
    A = matgen.laplace3d_7pt(96, 64, 40)
    x = matgen.vec_uniform(A.n, seed=6)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 4, x)
    dx = ctx.to_device(x)
    for flags in (0, 1, 2, 3):
        for pf in (1, 4):
            ctx.set_option("sell_flags", flags)
            ctx.set_option("sell_pf_dist", pf)
            lv = dA.mpk(4, dx)
            assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"flags={flags} pf={pf}")


def test_sell_splits_when_the_window_does_not_fit(sell, oracle_lib):
    ctx = sell
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
    assert 1 in seen and len(seen) >= 2, seen


@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (40,)), ("laplace2d_5pt", (300,)), ("tet_p1_laplacian", (20, 2, True)),
                                      ("fem_baij4", (6,))])
def test_sell_mpk_multi_bitwise(sell, oracle_lib, gen, args):
    """Two right-hand sides per launch (the s-step basis builder): every tile streamed once for both."""
    ctx = sell
    A = getattr(matgen, gen)(*args)
    xs = [matgen.vec_uniform(A.n, seed=20 + v) for v in range(3)]
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("wave_l2_pct", 1000)
    for k in (1, 2, 4):
        ref = [oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x) for x in xs]
        lv = dA.mpk_multi(k, [ctx.to_device(x) for x in xs])
        for v in range(3):
            for l in range(k):
                assert_bits_equal(lv[v][l].to_host(), ref[v][l], f"{gen} k={k} vector {v} level {l}")


def test_sell_cg_with_fused_dot(sell, oracle_lib):
    """CG through the sliced-ELL product (fused <p, Ap>): converges like the oracle, true residual checked."""
    ctx = sell
    A = matgen.laplace3d_7pt(48)
    xt = matgen.vec_uniform(A.n, seed=1)
    b = oracle_lib.spmv(A.ptrow, A.indcol, A.coef, xt)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    sol, it, rel, ok = dA.cg(b, tol=1e-8, maxit=500)
    assert ctx.query("last_spmv_kernel") == 4
    assert ok and oracle_lib.true_relres(A.ptrow, A.indcol, A.coef, b, sol) < 1e-7
    it_ref = oracle_lib.cg(A.ptrow, A.indcol, A.coef, b, 1e-8, 500)[1]
    assert abs(it - it_ref) <= 2


@pytest.mark.parametrize("seed", list(range(6)))
def test_sell_fuzz_random_stencils(sell, oracle_lib, seed):
    ctx = sell
    rng = np.random.default_rng(2000 + seed)
    nx, ny, nz = int(rng.integers(5, 70)), int(rng.integers(3, 40)), int(rng.integers(2, 24))
    A = matgen.random_stencil3d(nx, ny, nz, seed=seed, max_points=int(rng.integers(3, 14)), drop=float(rng.uniform(0, 0.2)))
    x = matgen.vec_uniform(A.n, seed=seed + 50)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("wave_l2_pct", 1000)
    ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, 4, x)
    assert_bits_equal(dA.spmv(x), ref[0], f"grid {nx}x{ny}x{nz}")
    dlv = dA.mpk(4, ctx.to_device(x))
    assert ctx.query("last_mpk_strategy") == 5
    assert_bits_equal(np.stack([l.to_host() for l in dlv]), ref, f"grid {nx}x{ny}x{nz} k=4")


def test_sell_ragged_and_empty_rows(sell, oracle_lib):
    ctx = sell
    for n, mean in [(1, 3.0), (7, 2.0), (1000, 0.5), (5000, 5.0), (3000, 40.0)]:
        A = matgen.random_csr(n, mean, seed=n, empty_rows=True)
        x = matgen.vec_uniform(n, seed=2)
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"n={n}")


@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (64, 64, 24)), ("laplace2d_5pt", (512, 300)), ("laplace3d_7pt", (33, 31, 40))])
@pytest.mark.parametrize("geom,tma", [(0, 0), (2, 0), (0, 4)])
def test_sell_long_streams_per_cta(sell, oracle_lib, gen, args, geom, tma):
    """A handful of CTAs walk the whole operator: every stage, ring slot and barrier phase is reused many times (straight-line
    and masked consumer paths, three and four stages)."""
    ctx = sell
    A = getattr(matgen, gen)(*args)
    x = matgen.vec_uniform(A.n, seed=3)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    ctx.set_option("sell_geom", geom)
    ctx.set_option("sell_tma", tma)
    ctx.set_option("wave_l2_pct", 1000)
    dx = ctx.to_device(x)
    try:
        for cap in (1, 3, 8, 29):
            ctx.set_option("sell_max_ctas", cap)
            assert_bits_equal(dA.spmv(x), oracle_lib.spmv(A.ptrow, A.indcol, A.coef, x), f"{gen}{args} spmv cap={cap}")
            assert ctx.query("last_spmv_kernel") == 4
            for k in (2, 4):
                ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
                for rep in range(2):
                    lv = dA.mpk(k, dx)
                    assert ctx.query("last_mpk_strategy") == 5
                    assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"{gen}{args} k={k} cap={cap} rep={rep}")
    finally:
        ctx.set_option("sell_max_ctas", 0)


def test_auto_strategy_fuses_unstructured_when_the_window_fits(ctx, oracle_lib):
    """Default options: an operator stored as explicit-column tiles (RCM-ordered tet P1 Laplacian) runs k >= 2 powers as ONE
    fused launch when min(k, 4) levels fit the L2 window, as k products otherwise (tiny L2 budget) or when the option
    mpk_auto_explicit is negative -- the same bits every way."""
    A = matgen.tet_p1_laplacian(24, 2, True)
    x = matgen.vec_uniform(A.n, seed=5)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    dx = ctx.to_device(x)
    try:
        lv = [ctx.zeros(A.n) for _ in range(2)]
        dA.mpk(2, dx, lv)
        assert ctx.query("last_mpk_strategy") == 1, "below the default size threshold (1 M rows): products"
        for k in (2, 4, 6):
            ctx.set_option("mpk_auto_explicit", 1)  # threshold: 1 row
            ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
            lv = [ctx.zeros(A.n) for _ in range(k)]
            before = ctx.launch_count
            dA.mpk(k, dx, lv)
            assert ctx.query("last_mpk_strategy") == 5 and ctx.launch_count - before == 1, f"k={k}: one fused launch expected"
            assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"auto fused k={k}")
            ctx.set_option("mpk_auto_explicit", -1)
            before = ctx.launch_count
            dA.mpk(k, dx, lv)
            assert ctx.query("last_mpk_strategy") == 1 and ctx.launch_count - before == k
            assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, f"auto off k={k}")
        ctx.set_option("mpk_auto_explicit", 1)
        ctx.set_option("wave_l2_pct", 1)  # no window fits 1 % of L2: products
        k = 4
        ref = oracle_lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
        lv = [ctx.zeros(A.n) for _ in range(k)]
        dA.mpk(k, dx, lv)
        assert ctx.query("last_mpk_strategy") == 1
        assert_bits_equal(np.stack([l.to_host() for l in lv]), ref, "window rejected")
    finally:
        ctx.set_option("mpk_auto_explicit", 0)
        ctx.set_option("wave_l2_pct", 0)
