"""Ingest path: COO -> CSR / 4x4 block CSR and the Matrix Market reader of libnsk.so against the
fixtures produced by the compiled reference and against the oracle's restatement.  Host-only: runs without a GPU."""
import numpy as np
import pytest

import navierstokes_b200 as nsk
from navierstokes_b200 import matgen

from conftest import assert_bits_equal, golden


def test_coo2csr_and_bcsr4_reference_fixture():
    g = golden("formats")
    n = int(g["n"])
    A = nsk.COO2CSR(nsk.csrmatrix(), n, g["irow"], g["jcol"], g["val"])
    assert A.n == n and A.nnz == len(g["irow"])  # nnz keeps the COO count (reference utils.cpp:100)
    assert np.array_equal(A.ptrow, g["csr_ptrow"]) and np.array_equal(A.indcol, g["csr_indcol"])
    assert_bits_equal(A.coef, g["csr_coef"])
    assert len(A.indcol) < len(g["irow"])  # the fixture contains duplicates: first wins
    B = nsk.generate_BCSR4(n, g["irow"], g["jcol"], g["val"])
    assert B.nblocks == 0  # the reference leaves nblocks at 0 (utils.cpp:78)
    assert np.array_equal(B.ptrow, g["bcsr_ptrow"]) and np.array_equal(B.indcol, g["bcsr_indcol"])
    assert_bits_equal(B.coef, g["bcsr_coef"])


@pytest.mark.parametrize("n,nnz,seed", [(8, 40, 1), (400, 9000, 2), (20000, 300000, 3), (4, 0, 4)])
def test_coo2csr_and_bcsr4_match_oracle_on_random_input(oracle_lib, n, nnz, seed):
    """Unsorted COO with many duplicates (including duplicates that differ in value) and empty rows."""
    rng = np.random.default_rng(seed)
    irow = rng.integers(0, n, nnz).astype(np.int32)
    jcol = rng.integers(0, n, nnz).astype(np.int32)
    if nnz:
        dup = rng.integers(0, nnz, nnz // 3)
        irow = np.concatenate([irow, irow[dup]])
        jcol = np.concatenate([jcol, jcol[dup]])
    val = rng.uniform(-1, 1, len(irow))
    A = nsk.COO2CSR(nsk.csrmatrix(), n, irow, jcol, val)
    p, c, v = oracle_lib.coo2csr(n, irow, jcol, val)
    assert np.array_equal(A.ptrow, p) and np.array_equal(A.indcol, c)
    assert_bits_equal(A.coef, v)
    B = nsk.generate_BCSR4(n, irow, jcol, val)
    bp, bc, bv = oracle_lib.generate_bcsr4(n, irow, jcol, val)
    assert np.array_equal(B.ptrow, bp) and np.array_equal(B.indcol, bc[:len(B.indcol)])
    assert_bits_equal(B.coef, bv[:len(B.coef)])


def test_coo2csr_rejects_bad_indices():
    with pytest.raises(nsk.NskError):
        nsk.COO2CSR(nsk.csrmatrix(), 4, [0, 5], [0, 1], [1.0, 2.0])


def test_mtx_reader_reference_quirks(tmp_path, oracle_lib):
    """Banner skipped unconditionally, % lines skipped, cols ignored, `symmetric` ignored, values through float32."""
    A = matgen.fem_baij4(2, float_round=False)
    rows = np.repeat(np.arange(A.n), np.diff(A.ptrow))
    path = tmp_path / "m.mtx"
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% a comment\n%another\n")
        f.write(f"{A.n} 999 {A.nnz}\n")
        for i, j, v in zip(rows, A.indcol, A.coef):
            f.write(f"{i + 1} {j + 1} {v:.17g}\n")
    n, irow, jcol, val = nsk.read_mtx(path)
    assert n == A.n and len(irow) == A.nnz
    assert np.array_equal(irow, rows) and np.array_equal(jcol, A.indcol)
    assert_bits_equal(val, A.coef.astype(np.float32).astype(np.float64))  # fp32-rounded like fscanf("%f")
    C = nsk.COO2CSR(nsk.csrmatrix(), n, irow, jcol, val)
    assert np.array_equal(C.ptrow, A.ptrow) and np.array_equal(C.indcol, A.indcol)
    with pytest.raises(nsk.NskError):
        nsk.read_mtx(tmp_path / "missing.mtx")


def test_ingest_is_fast_enough_for_large_inputs():
    """The reference's list-based build is O(nnz * row length); this one must do 10 M entries in seconds."""
    import time
    A = matgen.laplace3d_7pt(112)  # 1.4 M rows, 9.8 M entries
    rows = np.repeat(np.arange(A.n, dtype=np.int32), np.diff(A.ptrow))
    perm = np.random.default_rng(0).permutation(A.nnz)
    t0 = time.time()
    C = nsk.COO2CSR(nsk.csrmatrix(), A.n, rows[perm], A.indcol[perm], A.coef[perm])
    dt = time.time() - t0
    assert np.array_equal(C.ptrow, A.ptrow) and np.array_equal(C.indcol, A.indcol)
    assert_bits_equal(C.coef, A.coef)
    assert dt < 20.0, dt
