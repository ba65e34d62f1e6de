"""The packed format's host half (csrc/packed.cu pk_pack_host) without a GPU: pack, expand the blobs back to CSR
through the tiles' x runs and compare entry for entry with the input; refusals for operators that must not pack."""
import ctypes as C

import numpy as np
import pytest

from navierstokes_b200 import _lib, matgen

from conftest import CSR_CASES, assert_bits_equal, golden

NVARIANTS = 14


def pack(A_ptrow, A_indcol, A_coef, n_cols, variant):
    lib = _lib.load()
    ptrow = np.ascontiguousarray(A_ptrow, np.int32)
    indcol = np.ascontiguousarray(A_indcol, np.int32)
    coef = np.ascontiguousarray(A_coef, np.float64)
    n = len(ptrow) - 1
    h = C.c_void_p()
    st = lib.nsk_pack_host_create(n, n_cols, len(indcol), ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data, variant, C.byref(h))
    assert st == 0
    why = lib.nsk_pack_host_why(h).decode()
    out = None
    if not why:
        p2 = np.zeros(n + 1, np.int32)
        c2 = np.zeros(max(len(indcol), 1), np.int32)
        v2 = np.zeros(max(len(indcol), 1), np.float64)
        mr, mx = C.c_int(), C.c_int()
        assert lib.nsk_pack_host_expand(h, p2.ctypes.data, c2.ctypes.data, v2.ctypes.data, C.byref(mr), C.byref(mx)) == 0
        out = (p2, c2[:len(indcol)], v2[:len(indcol)], mr.value, mx.value, int(lib.nsk_pack_host_bytes(h)))
    lib.nsk_pack_host_destroy(h)
    return why, out


@pytest.mark.parametrize("variant", range(NVARIANTS))
@pytest.mark.parametrize("gen,args", [("laplace3d_7pt", (24, 18, 10)), ("laplace2d_5pt", (130, 41)), ("laplace3d_7pt", (300, 4, 6)),
                                      ("fem_baij4", (4,)), ("laplace3d_7pt", (11, 9, 7)), ("laplace2d_5pt", (33, 29)),
                                      ("laplace3d_7pt", (257, 3, 1))])
def test_pack_expand_round_trip(gen, args, variant):
    A = getattr(matgen, gen)(*args)
    why, out = pack(A.ptrow, A.indcol, A.coef, A.n, variant)
    if why:  # a geometry whose stage is too small for this operator's rows / runs may refuse; it must say why
        assert "row" in why or "runs" in why
        return
    p2, c2, v2, max_runs, max_xlen, nbytes = out
    assert np.array_equal(p2, A.ptrow) and np.array_equal(c2, A.indcol)
    assert_bits_equal(v2, A.coef)
    assert 1 <= max_runs <= 8 and nbytes > 0


def test_pack_stencils_use_few_runs_and_ten_bytes_per_nonzero():
    A = matgen.laplace3d_7pt(64, 64, 16)
    why, out = pack(A.ptrow, A.indcol, A.coef, A.n, 7)
    assert not why
    _, _, _, max_runs, max_xlen, nbytes = out
    assert max_runs == 3                      # plane below, own lines +-1, plane above
    assert max_xlen <= 256 + 256 + 2 * 64 + 256 + 16
    assert nbytes / A.nnz < 10.6              # 8 (value) + 2 (local column) + lens/header/padding


@pytest.mark.parametrize("seed", range(6))
def test_pack_random_stencils_round_trip(seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = int(rng.integers(4, 50)), int(rng.integers(3, 30)), int(rng.integers(2, 16))
    A = matgen.random_stencil3d(nx, ny, nz, seed=seed, max_points=int(rng.integers(3, 14)), drop=float(rng.uniform(0, 0.2)))
    why, out = pack(A.ptrow, A.indcol, A.coef, A.n, 7)
    if why:
        return
    p2, c2, v2, *_ = out
    assert np.array_equal(p2, A.ptrow) and np.array_equal(c2, A.indcol)
    assert_bits_equal(v2, A.coef)


def test_pack_refusals():
    rnd = matgen.random_csr(4000, 5.0, seed=1)
    why, _ = pack(rnd.ptrow, rnd.indcol, rnd.coef, rnd.n, 7)
    assert why  # scattered columns: too many runs (or too ragged)
    ragged = matgen.random_banded_csr(20000, 150, 5.0, seed=4)
    why, _ = pack(ragged.ptrow, ragged.indcol, ragged.coef, ragged.n, 7)
    assert "ragged" in why
    tet = matgen.tet_p1_laplacian(30, permute_seed=2, rcm=True)  # RCM band of a few thousand columns: too many runs
    why, _ = pack(tet.ptrow, tet.indcol, tet.coef, tet.n, 7)
    assert why


@pytest.mark.parametrize("case", CSR_CASES)
def test_pack_golden_operators(case):
    """The fixtures produced by the compiled reference: whatever packs expands back to exactly the fixture."""
    g = golden(case)
    n = len(g["ptrow"]) - 1
    for variant in (7, 10):
        why, out = pack(g["ptrow"], g["indcol"], g["coef"], n, variant)
        if why:
            continue
        p2, c2, v2, *_ = out
        assert np.array_equal(p2, g["ptrow"]) and np.array_equal(c2, g["indcol"])
        assert_bits_equal(v2, g["coef"])


# ---- the fused kernel's protocol, modelled on the CPU ----------------------------------------------------------------
def simulate(A, variant, k, slack, resident, w0=100, bp_global=1, interleave=1, stages=2, level_rows=None, seed=0,
             ghi_bias=0):
    lib = _lib.load()
    ptrow = np.ascontiguousarray(A.ptrow, np.int32)
    indcol = np.ascontiguousarray(A.indcol, np.int32)
    coef = np.ascontiguousarray(A.coef, np.float64)
    h = C.c_void_p()
    assert lib.nsk_pack_host_create(A.n, A.n, len(indcol), ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data, variant,
                                    C.byref(h)) == 0
    if lib.nsk_pack_host_why(h):
        lib.nsk_pack_host_destroy(h)
        return None
    lr = None
    if level_rows is not None:
        lr = np.ascontiguousarray(level_rows, np.int32)
    items, reach = C.c_longlong(), C.c_int()
    stuck = lib.nsk_pack_host_simulate(h, k, slack, resident, w0, bp_global, interleave, stages,
                                       lr.ctypes.data if lr is not None else None, seed, C.byref(items), C.byref(reach),
                                       ghi_bias)
    lib.nsk_pack_host_destroy(h)
    return stuck, items.value, reach.value


@pytest.mark.parametrize("bp_global", [0, 1])
@pytest.mark.parametrize("k", [1, 2, 4, 7, 16])
@pytest.mark.parametrize("slack", [0, 5, 300])
def test_protocol_model_never_deadlocks(k, slack, bp_global):
    """Forward dependencies + window back-pressure on the schedule the GPU path builds: every item becomes runnable,
    from the tightest window (lead = reach + one completion group) to a loose one, few or many CTAs, even or uneven
    teams, both placements, 1-3 open items per CTA, several random interleavings."""
    A = matgen.laplace3d_7pt(64, 24, 20)  # 120 tiles of 256 rows, reach 6 tiles
    rng = np.random.default_rng(k * 100 + slack + bp_global)
    for trial in range(6):
        resident = int(rng.choice([k, k + 1, 3 * k, 40, 444]))
        if resident < k:
            resident = k
        w0 = int(rng.choice([100, 40, 300]))
        stuck, items, reach = simulate(A, 7, k, slack, resident, w0=w0, bp_global=bp_global,
                                       interleave=int(rng.integers(0, 2)), stages=int(rng.integers(1, 4)), seed=trial)
        assert stuck == 0, (k, slack, resident, w0, stuck, items)
        assert items == k * 120 and 6 <= reach <= 6 + 16


def test_protocol_model_with_shrinking_row_prefixes():
    """Distributed slabs evaluate level l on a row prefix (ghost rings drop out level by level): groups that lose all
    their tiles must count as complete, partially covered groups must expect fewer reports."""
    A = matgen.laplace3d_7pt(64, 16, 40)  # planes of 1024 rows = 4 tiles
    k = 4
    lr = [A.n - 1024 * l for l in range(k)]  # one plane less per level
    for seed in range(4):
        stuck, items, _ = simulate(A, 7, k, 3, 37, level_rows=lr, seed=seed, stages=2)
        assert stuck == 0
        assert items == sum(r // 256 for r in lr)


def test_protocol_model_other_patterns():
    ran = 0
    ops = [matgen.laplace2d_5pt(300, 40), matgen.fem_baij4(5)] + [matgen.random_stencil3d(30, 20, 12, seed=s, max_points=6)
                                                                  for s in range(6)]
    for A in ops:
        for k in (2, 5):
            res = simulate(A, 10 if A.nnz / A.n > 16 else 7, k, 2, 50, seed=1)
            if res is None:  # this random stencil needs more than 8 runs per tile: it does not pack
                continue
            stuck, items, _ = res
            assert stuck == 0 and items > 0
            ran += 1
    assert ran >= 6


def test_protocol_model_detects_a_window_smaller_than_the_reach():
    """The model has teeth: a window that lets level l lead level l+1 by less than the pattern's reach must deadlock
    (level l+1 waits for tiles level l is not allowed to start) -- which is why the planner never goes below
    reach + one group."""
    A = matgen.laplace3d_7pt(64, 24, 20)
    stuck, items, reach = simulate(A, 7, 3, -(6 + 16 + 1) - 10, 30, bp_global=0, seed=0)
    assert stuck > 0 and stuck < items
    stuck, _, _ = simulate(A, 7, 3, -(6 + 16 + 1) - 10, 30, bp_global=1, seed=0)
    assert stuck > 0
    stuck, _, _ = simulate(A, 7, 3, 0, 30, bp_global=0, seed=0)
    assert stuck == 0


def test_protocol_model_detects_missing_forward_dependencies():
    """Data readiness: with the forward dependencies weakened by a few groups, some item opens before a tile it really
    reads is complete at the level below -- the model reports the hole (on the real schedule it never does: every
    other test in this file runs with the check on)."""
    A = matgen.laplace3d_7pt(64, 24, 20)  # reach 6 tiles: one group back is enough to lose the plane above
    found = 0
    for seed in range(8):
        res, _, _ = simulate(A, 7, 3, 40, 444, seed=seed, ghi_bias=-2)
        found += res <= -1000000
    assert found >= 1
    res, _, _ = simulate(A, 7, 3, 40, 444, seed=0, ghi_bias=0)
    assert res == 0
