"""The sliced-ELL format's host half (csrc/sell.cu) without a GPU: tiling, pattern detection, blobs (expanded back to
CSR and compared entry for entry with the input), and the fused kernel's protocol modelled on the CPU on the very
schedule the GPU path builds."""
import ctypes as C

import numpy as np
import pytest

from navierstokes_b200 import _lib, matgen

from conftest import CSR_CASES, assert_bits_equal, golden


def sell(ptrow, indcol, coef, n_cols):
    lib = _lib.load()
    ptrow = np.ascontiguousarray(ptrow, np.int32)
    indcol = np.ascontiguousarray(indcol, np.int32)
    coef = np.ascontiguousarray(coef, np.float64)
    n = len(ptrow) - 1
    h = C.c_void_p()
    assert lib.nsk_sell_host_create(n, n_cols, len(indcol), ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data,
                                    C.byref(h)) == 0
    why = lib.nsk_sell_host_why(h).decode()
    out = None
    if not why:
        p2 = np.zeros(n + 1, np.int32)
        c2 = np.zeros(max(len(indcol), 1), np.int32)
        v2 = np.zeros(max(len(indcol), 1), np.float64)
        assert lib.nsk_sell_host_expand(h, p2.ctypes.data, c2.ctypes.data, v2.ctypes.data) == 0
        b, nt, npat = C.c_int64(), C.c_int64(), C.c_int64()
        assert lib.nsk_sell_host_stats(h, C.byref(b), C.byref(nt), C.byref(npat)) == 0
        out = (p2, c2[:len(indcol)], v2[:len(indcol)], b.value, nt.value, npat.value)
    lib.nsk_sell_host_destroy(h)
    return why, out


OPS = [("laplace3d_7pt", (24, 18, 10)), ("laplace2d_5pt", (130, 41)), ("laplace3d_7pt", (300, 4, 6)), ("fem_baij4", (4,)),
       ("laplace3d_7pt", (11, 9, 7)), ("laplace2d_5pt", (33, 29)), ("laplace3d_7pt", (257, 3, 1)), ("laplace3d_7pt", (256, 6, 5)),
       ("tet_p1_laplacian", (12,)), ("random_csr", (3000, 5.0)), ("random_banded_csr", (5000, 150, 9.0))]


@pytest.mark.parametrize("gen,args", OPS)
def test_sell_expand_round_trip(gen, args):
    A = getattr(matgen, gen)(*args)
    why, out = sell(A.ptrow, A.indcol, A.coef, A.n)
    if why:
        assert "ragged" in why
        return
    p2, c2, v2, nbytes, ntiles, npat = out
    assert np.array_equal(p2, A.ptrow) and np.array_equal(c2, A.indcol)
    assert_bits_equal(v2, A.coef)
    assert ntiles == (A.n + 255) // 256 and 0 <= npat <= ntiles and nbytes > 0


def test_sell_stencils_are_pattern_tiles_at_eight_bytes_per_nonzero():
    """7-point operator, 256-row tiles = x-lines: every tile has one column pattern; its two line-end rows lack one slot
    each, lines on the y / z faces lack a whole neighbour line and have a narrower pattern.  No per-entry index."""
    A = matgen.laplace3d_7pt(256, 6, 5)
    why, out = sell(A.ptrow, A.indcol, A.coef, A.n)
    assert not why
    *_, nbytes, ntiles, npat = out
    assert npat == ntiles == 30
    assert nbytes / A.nnz < 9.9  # 8 bytes per slot, every tile padded to the operator's 7 slots (most lines of this tiny
                                 # grid lie on a face and use 5 or 6 of them)
    big = matgen.laplace3d_7pt(256, 24, 20)
    _, out_big = sell(big.ptrow, big.indcol, big.coef, big.n)
    assert out_big[3] / big.nnz < 8.75
    B = matgen.laplace3d_7pt(61, 17, 23)  # tiles straddle lines: still one pattern per tile (rel = the 7 stencil offsets)
    why, out = sell(B.ptrow, B.indcol, B.coef, B.n)
    assert not why
    *_, nbytes, ntiles, npat = out
    assert npat == ntiles
    C5 = matgen.laplace2d_5pt(1024, 7)
    why, out = sell(C5.ptrow, C5.indcol, C5.coef, C5.n)
    assert not why and out[5] == out[4]


def test_sell_unstructured_operators_keep_explicit_columns():
    tet = matgen.tet_p1_laplacian(20, permute_seed=2, rcm=True)
    why, out = sell(tet.ptrow, tet.indcol, tet.coef, tet.n)
    assert not why
    p2, c2, v2, nbytes, ntiles, npat = out
    assert np.array_equal(c2, tet.indcol) and npat < ntiles
    assert nbytes < 1.5 * (12 * tet.nnz + 4 * tet.n) + 65536
    fem = matgen.fem_baij4(5)   # 58 per row: explicit tiles with per-slice widths
    why, out = sell(fem.ptrow, fem.indcol, fem.coef, fem.n)
    assert not why and np.array_equal(out[1], fem.indcol)


def test_sell_refuses_very_ragged_rows():
    rows = 4096
    lens = np.ones(rows, np.int64)
    lens[::32] = 200  # one long row per 32-row slice: slot-major slices would be almost all padding
    ptrow = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    indcol = np.concatenate([np.arange(l) for l in lens]).astype(np.int32)
    coef = np.ones(len(indcol))
    why, _ = sell(ptrow, indcol, coef, rows)
    assert "ragged" in why


@pytest.mark.parametrize("seed", range(6))
def test_sell_random_stencils_round_trip(seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = int(rng.integers(4, 50)), int(rng.integers(3, 30)), int(rng.integers(2, 16))
    A = matgen.random_stencil3d(nx, ny, nz, seed=seed, max_points=int(rng.integers(3, 14)), drop=float(rng.uniform(0, 0.2)))
    why, out = sell(A.ptrow, A.indcol, A.coef, A.n)
    assert not why
    p2, c2, v2, *_ = out
    assert np.array_equal(p2, A.ptrow) and np.array_equal(c2, A.indcol)
    assert_bits_equal(v2, A.coef)


@pytest.mark.parametrize("case", CSR_CASES)
def test_sell_golden_operators(case):
    """The fixtures produced by the compiled reference expand back to exactly the fixture (empty rows, ragged rows,
    unsorted columns, duplicates included)."""
    g = golden(case)
    n = len(g["ptrow"]) - 1
    why, out = sell(g["ptrow"], g["indcol"], g["coef"], n)
    if why:
        assert "ragged" in why or "empty" in why
        return
    p2, c2, v2, *_ = out
    assert np.array_equal(p2, g["ptrow"]) and np.array_equal(c2, g["indcol"])
    assert_bits_equal(v2, g["coef"])


# ---- the fused kernel's protocol, modelled on the CPU ----------------------------------------------------------------
def simulate(A, k, chunk, slack, resident, w0=100, interleave=1, ring=4, level_rows=None, seed=0, pmax_bias=0):
    lib = _lib.load()
    ptrow = np.ascontiguousarray(A.ptrow, np.int32)
    indcol = np.ascontiguousarray(A.indcol, np.int32)
    coef = np.ascontiguousarray(A.coef, np.float64)
    h = C.c_void_p()
    assert lib.nsk_sell_host_create(A.n, A.n, len(indcol), ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data,
                                    C.byref(h)) == 0
    if lib.nsk_sell_host_why(h):
        lib.nsk_sell_host_destroy(h)
        return None
    lr = None
    if level_rows is not None:
        lr = np.ascontiguousarray(level_rows, np.int32)
    items, reach = C.c_longlong(), C.c_int()
    stuck = lib.nsk_sell_host_simulate(h, k, chunk, slack, resident, w0, interleave, ring,
                                       lr.ctypes.data if lr is not None else None, seed, C.byref(items), C.byref(reach),
                                       pmax_bias)
    lib.nsk_sell_host_destroy(h)
    return stuck, items.value, reach.value


@pytest.mark.parametrize("chunk", [1, 2, 3, 4])
@pytest.mark.parametrize("k", [1, 2, 4, 7, 16])
@pytest.mark.parametrize("slack", [0, 5, 300])
def test_protocol_model_never_deadlocks(k, slack, chunk):
    """Forward dependencies + window back-pressure on the schedule the GPU path builds: every item is published, from the
    tightest window to a loose one, few or many CTAs, even or uneven teams, both placements, 1-4 open items per CTA."""
    A = matgen.laplace3d_7pt(64, 24, 20)  # 120 tiles of 256 rows, reach 6 tiles
    rng = np.random.default_rng(k * 100 + slack + chunk)
    for trial in range(5):
        resident = max(k, int(rng.choice([k, k + 1, 3 * k, 40, 444])))
        w0 = int(rng.choice([100, 40, 300]))
        stuck, items, reach = simulate(A, k, chunk, slack, resident, w0=w0, interleave=int(rng.integers(0, 2)),
                                       ring=int(rng.integers(1, 5)), seed=trial)
        assert stuck == 0, (k, slack, chunk, resident, w0, stuck, items)
        assert items == k * ((120 + chunk - 1) // chunk) and reach == 6


def test_protocol_model_with_shrinking_row_prefixes():
    """Distributed slabs evaluate level l on a row prefix: each level has its own tile list, groups that lose all their
    items count as complete, dependencies are looked up in the level below's own list."""
    A = matgen.laplace3d_7pt(64, 16, 40)  # planes of 1024 rows = 4 tiles
    k = 4
    lr = [A.n - 1024 * l for l in range(k)]
    for chunk in (1, 2, 3):
        for seed in range(3):
            stuck, items, _ = simulate(A, k, chunk, 3, 37, level_rows=lr, seed=seed)
            assert stuck == 0
            assert items == sum((r // 256 + chunk - 1) // chunk for r in lr)


def test_protocol_model_other_patterns():
    ops = [matgen.laplace2d_5pt(300, 40), matgen.fem_baij4(5), matgen.tet_p1_laplacian(14, permute_seed=2, rcm=True)]
    ops += [matgen.random_stencil3d(30, 20, 12, seed=s, max_points=6) for s in range(4)]
    for A in ops:
        for k in (2, 5):
            stuck, items, _ = simulate(A, k, 2, 2, 50, seed=1)
            assert stuck == 0 and items > 0


def test_protocol_model_detects_a_window_smaller_than_the_reach():
    """The model has teeth: a window that lets level 0 lead level k-1 by less than the pattern's reach must deadlock."""
    A = matgen.laplace3d_7pt(64, 24, 20)
    stuck, items, reach = simulate(A, 3, 1, -(6 + 16 + 1 + 1) - 12, 30, seed=0)
    assert 0 < stuck < items
    stuck, _, _ = simulate(A, 3, 1, 0, 30, seed=0)
    assert stuck == 0


def test_protocol_model_detects_missing_forward_dependencies():
    """Data readiness: with the forward dependencies weakened, some item opens before a tile it really reads is
    published at the level below -- the model reports the hole (on the real schedule it never does: every other test in
    this file runs with the check on)."""
    A = matgen.laplace3d_7pt(64, 24, 20)
    found = 0
    for seed in range(8):
        res, _, _ = simulate(A, 3, 1, 40, 444, seed=seed, pmax_bias=-40)
        found += res <= -1000000
    assert found >= 1
    res, _, _ = simulate(A, 3, 1, 40, 444, seed=0, pmax_bias=0)
    assert res == 0


def _global_pattern(ptrow, indcol, coef, n_cols):
    lib = _lib.load()
    ptrow = np.ascontiguousarray(ptrow, np.int32)
    indcol = np.ascontiguousarray(indcol, np.int32)
    coef = np.ascontiguousarray(coef, np.float64)
    h = C.c_void_p()
    assert lib.nsk_sell_host_create(len(ptrow) - 1, n_cols, len(indcol), ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data,
                                    C.byref(h)) == 0
    rel = np.zeros(8, np.int32)
    g, nt = C.c_int64(0), C.c_int64(0)
    w = lib.nsk_sell_host_global_pattern(h, rel.ctypes.data, C.byref(g))
    lib.nsk_sell_host_stats(h, None, C.byref(nt), None)
    lib.nsk_sell_host_destroy(h)
    return w, rel, g.value, nt.value


def test_sell_global_pattern_of_a_stencil():
    """A 7-point operator on a box has ONE pattern for all tiles (faces only lack offsets): the kernel's straight-line path."""
    A = matgen.laplace3d_7pt(256, 6, 5)
    w, rel, g, nt = _global_pattern(A.ptrow, A.indcol, A.coef, A.n)
    assert w == 7 and g == nt
    assert list(rel[:7]) == [-256 * 6, -256, -1, 0, 1, 256, 256 * 6]


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_sell_global_pattern_of_a_distributed_slab_has_exception_tiles(rank):
    """The local operator of a z-slab with depth-4 ghost rings: the tiles next to a ring reference it at offsets of their own
    and keep their own pattern (exceptions); everything else shares the global pattern; the blobs still expand to the input."""
    from navierstokes_b200 import distributed as nd
    nx, ny, nz = 256, 5, 12
    plan = nd.Plan.build(3, rank, nd.slab_row_starts(nz, nx * ny, 3), 4, nd.StencilProvider(nx, ny, nz))
    A = plan.local_csr()
    why, out = sell(A.ptrow, A.indcol, A.coef, A.ncols)
    assert not why
    p2, c2, v2, _, ntiles, npat = out
    assert np.array_equal(p2, A.ptrow) and np.array_equal(c2, A.indcol)
    assert_bits_equal(v2, A.coef)
    assert npat == ntiles
    w, rel, g, nt = _global_pattern(A.ptrow, A.indcol, A.coef, A.ncols)
    # rank 0 only has rings above its slab, stored right behind it at the stencil's own offsets: no exception needed
    assert w == 7 and nt == ntiles and 10 <= g <= nt and (g < nt or rank == 0), (w, g, nt)
