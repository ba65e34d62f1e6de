"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled in oracle/_ref.

Run in the development container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle.py) and the CUDA path (tests/test_*_gpu.py) to outputs
of the reference's own code: SpMV_CSR{,_OPT,_FMA,_AVX2}, SpM2V_CSR{,_OPT}, SpM2V0/SpM3V/SpM4V with
their Generate*layer schedules, COO2CSR, generate_BCSR4, SpMV_BCSR*, norm2, rel_error, orthogonalize, orthonormalize_against_basis.
The reference itself ships no vectors or matrices, so these are the golden data.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402

OUT = Path(__file__).resolve().parent


def vectors(n):
    return {"ones": matgen.vec_ones(n), "sin": matgen.vec_sin(n), "uni": matgen.vec_uniform(n, seed=1)}


def csr_case(name, A, deep=False):
    oracle.build()
    ref = oracle.ref
    R = ref.csr(A.ptrow, A.indcol, A.coef)
    d = {"ptrow": A.ptrow, "indcol": A.indcol, "coef": A.coef}
    mult4 = bool(np.all(np.diff(A.ptrow) % 4 == 0))
    d["ptrowend1"] = R.generate_1st_layer()
    for vn, x in vectors(A.n).items():
        d[f"x_{vn}"] = x
        for var in ("x87", "opt", "fma") + (("avx2",) if mult4 else ()):
            d[f"spmv_{var}_{vn}"] = R.spmv(x, var)
        for var in ("x87", "opt") + (("avx2",) if mult4 else ()):
            y, z = R.spm2v(x, var, d["ptrowend1"])
            d[f"spm2v_{var}_y_{vn}"] = y
            d[f"spm2v_{var}_z_{vn}"] = z
        if deep:
            for depth in (2, 3, 4):
                d[f"multi0_k{depth}_{vn}"] = ref.multi0_spmkv(A.ptrow, A.indcol, A.coef, depth, x)
            d[f"multi0_spmv_{vn}"] = ref.multi0_spmv(A.ptrow, A.indcol, A.coef, x)
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, A.nrows, A.nnz, "mult4" if mult4 else "")


def formats_case():
    ref = oracle.ref
    rng = np.random.default_rng(7)
    n, nnz = 96, 1500
    irow = rng.integers(0, n, nnz).astype(np.int32)
    jcol = rng.integers(0, n, nnz).astype(np.int32)
    val = rng.uniform(-1, 1, nnz).astype(np.float32).astype(np.float64)  # reader rounds through float
    d = {"n": n, "irow": irow, "jcol": jcol, "val": val}
    d["csr_ptrow"], d["csr_indcol"], d["csr_coef"] = ref.coo2csr(n, irow, jcol, val)
    d["bcsr_ptrow"], d["bcsr_indcol"], d["bcsr_coef"] = ref.generate_bcsr4(n, irow, jcol, val)
    x = rng.uniform(-1, 1, n)
    d["x"] = x
    for var in ("x87", "opt", "fma", "avx2"):
        d[f"bcsr_spmv_{var}"] = ref.spmv_bcsr4(d["bcsr_ptrow"], d["bcsr_indcol"], d["bcsr_coef"], x, var)
    for var in ("x87", "opt", "avx2"):
        y, z = ref.spm2v_bcsr4(d["bcsr_ptrow"], d["bcsr_indcol"], d["bcsr_coef"], x, var)
        d[f"bcsr_spm2v_{var}_y"] = y
        d[f"bcsr_spm2v_{var}_z"] = z
    a = rng.uniform(-1, 1, 1003)
    b = a + 1e-9 * rng.uniform(-1, 1, 1003)
    c = rng.uniform(-1, 1, 1003)
    d["va"], d["vb"], d["vc"] = a, b, c
    d["norm2_a"] = ref.norm2(a)
    d["rel_error_ab"] = ref.rel_error(a, b)
    d["orth_ac"] = ref.orthogonalize(a, c, 1e-8)
    d["orth_ac_alpha1"] = ref.orthogonalize(a, c, 0.37)
    np.savez_compressed(OUT / "formats.npz", **d)
    print("formats", len(d["csr_indcol"]), "kept of", nnz, "; bcsr blocks", len(d["bcsr_indcol"]))


def orthobasis_case():
    """orthonormalize_against_basis (mpk/2SpMV.cpp:13-28) on the reference's own fake Krylov basis sin(0.001 j + i)
    (mpk/2SpMV.cpp:110-116), shortened to 6 vectors of 1003."""
    ref = oracle.ref
    n, m = 1003, 6
    basis = [np.sin(0.001 * j + np.arange(n)) for j in range(m)]
    y = np.random.default_rng(11).uniform(-1, 1, n)
    d = {"basis": np.stack(basis), "y": y, "y_out": ref.orthonormalize_against_basis(basis, y)}
    np.savez_compressed(OUT / "orthobasis.npz", **d)
    print("orthobasis", n, m)


if __name__ == "__main__":
    assert oracle.REFERENCE_SRC.is_dir(), "needs /root/reference (development container)"
    if "--only-orthobasis" in sys.argv:
        oracle.build()
        orthobasis_case()
        sys.exit(0)
    csr_case("lap3d_7pt_6", matgen.laplace3d_7pt(6), deep=True)
    csr_case("lap3d_7pt_12x10x9", matgen.laplace3d_7pt(12, 10, 9))
    csr_case("lap2d_5pt_33x29", matgen.laplace2d_5pt(33, 29))
    csr_case("fem_baij4_m3", matgen.fem_baij4(3))
    csr_case("fem_baij4_m2_deep", matgen.fem_baij4(2), deep=True)
    csr_case("tet_p1_m5_rcm", matgen.tet_p1_laplacian(5, permute_seed=2, rcm=True), deep=True)
    csr_case("ragged_300", matgen.random_csr(300, 6.0, seed=3, empty_rows=True))
    formats_case()
    orthobasis_case()
