"""CPU oracle for the SpMV / matrix-powers / CG hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under navierstokes_b200/ does.

Two checkers live here:

* ``oracle.lib``  -- ctypes binding of oracle/liboracle.so, our plain-C restatement of the
  reference algorithms (oracle/nsk_oracle.c, each function cites the reference file:line).
* ``oracle.ref``  -- ctypes binding of oracle/_ref/libnsref_*.so, the UNMODIFIED reference
  sources compiled from /root/reference/mpk (oracle/Makefile).  Present only when it was built
  in the development container; the prebuilt files travel to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REFERENCE_SRC = Path("/root/reference/mpk")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref: bool | None = None) -> None:
    """Compile liboracle.so and, when the reference sources are present, oracle/_ref/*.so."""
    subprocess.run(["make", "-s", "-C", str(_HERE), "oracle"], check=True)
    if ref is None:
        ref = REFERENCE_SRC.is_dir()
    if ref:
        subprocess.run(["make", "-s", "-C", str(_HERE), "ref"], check=True)


def _load(path: Path) -> C.CDLL:
    if not path.exists():
        raise FileNotFoundError(f"{path} is not built; run `make -C oracle` (see oracle/Makefile)")
    return C.CDLL(str(path))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Oracle:
    """Binding of oracle/liboracle.so (the C restatement)."""

    def __init__(self):
        self._l = None

    @property
    def l(self):
        if self._l is None:
            so = _HERE / "liboracle.so"
            if not so.exists():
                build(ref=False)
            l = _load(so)
            l.oracle_spmv_csr_fma.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_spmv_csr_fma_rows.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_spmv_csr_muladd.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_spmv_csr_avx2.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_spmv_csr_avx2.restype = C.c_int
            l.oracle_mpk_csr_fma.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p]
            l.oracle_generate_1st_layer.argtypes = [C.c_int, _i32p, _i32p, _i32p]
            l.oracle_spm2v_csr.argtypes = [C.c_int, _i32p, _i32p, _f64p, _i32p, C.c_int, C.c_int, _f64p, _f64p, _f64p]
            l.oracle_spmkv_first_touch.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_int, _f64p, _f64p]
            l.oracle_spmkv_first_touch.restype = C.c_int
            l.oracle_coo2csr.argtypes = [C.c_int, C.c_int64, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
            l.oracle_coo2csr.restype = C.c_int64
            l.oracle_generate_bcsr4.argtypes = [C.c_int, C.c_int64, _i32p, _i32p, _f64p,
                                                C.c_void_p, C.c_void_p, C.c_void_p]
            l.oracle_generate_bcsr4.restype = C.c_int64
            l.oracle_spmv_bcsr4_fma.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_spmm_baij4.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                            C.c_longlong, C.c_int]
            l.oracle_norm2.argtypes = [C.c_int64, _f64p]
            l.oracle_norm2.restype = C.c_double
            l.oracle_rel_error.argtypes = [C.c_int64, _f64p, _f64p]
            l.oracle_rel_error.restype = C.c_double
            l.oracle_dot.argtypes = [C.c_int64, _f64p, _f64p]
            l.oracle_dot.restype = C.c_double
            l.oracle_orthogonalize.argtypes = [C.c_int64, _f64p, _f64p, C.c_double]
            l.oracle_orthogonalize.restype = C.c_double
            l.oracle_orthonormalize_against_basis.argtypes = [C.c_int64, C.c_int, _f64p, _f64p]
            l.oracle_orthonormalize_against_basis.restype = C.c_double
            l.oracle_cg.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_double, C.c_int,
                                    C.POINTER(C.c_double), C.c_void_p]
            l.oracle_cg.restype = C.c_int
            l.oracle_true_relres.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            l.oracle_true_relres.restype = C.c_double
            self._l = l
        return self._l

    # -- SpMV ------------------------------------------------------------------------------
    def spmv(self, ptrow, indcol, coef, x):
        n = len(ptrow) - 1
        y = np.empty(n, dtype=np.float64)
        self.l.oracle_spmv_csr_fma(n, _i32(ptrow), _i32(indcol), _f64(coef), _f64(x), y)
        return y

    def spmv_rows(self, r0, r1, ptrow, indcol, coef, x, y):
        """In-place y[r0:r1] = (A x)[r0:r1]; arrays must already be contiguous of the right dtype."""
        self.l.oracle_spmv_csr_fma_rows(r0, r1, ptrow, indcol, coef, x, y)

    def spmv_muladd(self, ptrow, indcol, coef, x):
        n = len(ptrow) - 1
        y = np.empty(n, dtype=np.float64)
        self.l.oracle_spmv_csr_muladd(n, _i32(ptrow), _i32(indcol), _f64(coef), _f64(x), y)
        return y

    def spmv_avx2(self, ptrow, indcol, coef, x):
        n = len(ptrow) - 1
        y = np.empty(n, dtype=np.float64)
        rc = self.l.oracle_spmv_csr_avx2(n, _i32(ptrow), _i32(indcol), _f64(coef), _f64(x), y)
        if rc:
            raise ValueError("row length not a multiple of 4: the reference AVX2 kernel is undefined")
        return y

    def mpk(self, ptrow, indcol, coef, k, x):
        n = len(ptrow) - 1
        out = np.empty((k, n), dtype=np.float64)
        self.l.oracle_mpk_csr_fma(n, _i32(ptrow), _i32(indcol), _f64(coef), k, _f64(x), out.reshape(-1))
        return out

    def generate_1st_layer(self, ptrow, indcol):
        n = len(ptrow) - 1
        pe = np.empty(len(indcol), dtype=np.int32)
        self.l.oracle_generate_1st_layer(n, _i32(ptrow), _i32(indcol), pe)
        return pe

    def spm2v(self, ptrow, indcol, coef, x, ptrowend1=None, y_fma=True, z_fma=True):
        """Lazy first-touch A^2 x; (y_fma, z_fma) = (False, True) is the reference's _OPT as built here."""
        n = len(ptrow) - 1
        if ptrowend1 is None:
            ptrowend1 = self.generate_1st_layer(ptrow, indcol)
        y = np.empty(n)
        z = np.empty(n)
        self.l.oracle_spm2v_csr(n, _i32(ptrow), _i32(indcol), _f64(coef), _i32(ptrowend1), int(y_fma), int(z_fma),
                                _f64(x), y, z)
        return y, z

    def spmkv_first_touch(self, ptrow, indcol, coef, depth, x, use_fma=True):
        n = len(ptrow) - 1
        out = np.zeros((4, n), dtype=np.float64)
        rc = self.l.oracle_spmkv_first_touch(n, _i32(ptrow), _i32(indcol), _f64(coef), depth,
                                             1 if use_fma else 0, _f64(x), out.reshape(-1))
        if rc:
            raise ValueError("depth must be 2, 3 or 4")
        return out[:depth]

    # -- formats ---------------------------------------------------------------------------
    def coo2csr(self, nrow, irow, jcol, val):
        nnz = len(irow)
        ptrow = np.zeros(nrow + 1, dtype=np.int32)
        indcol = np.zeros(max(nnz, 1), dtype=np.int32)
        coef = np.zeros(max(nnz, 1), dtype=np.float64)
        kept = self.l.oracle_coo2csr(nrow, nnz, _i32(irow), _i32(jcol), _f64(val), ptrow, indcol, coef)
        return ptrow, indcol[:kept].copy(), coef[:kept].copy()

    def generate_bcsr4(self, nrow, irow, jcol, val):
        nnz = len(irow)
        irow, jcol, val = _i32(irow), _i32(jcol), _f64(val)
        nblk = self.l.oracle_generate_bcsr4(nrow, nnz, irow, jcol, val, None, None, None)
        ptrow = np.zeros(nrow // 4 + 1, dtype=np.int32)
        indcol = np.zeros(max(nblk, 1), dtype=np.int32)
        coef = np.zeros(16 * max(nblk, 1), dtype=np.float64)
        self.l.oracle_generate_bcsr4(nrow, nnz, irow, jcol, val, ptrow.ctypes.data, indcol.ctypes.data,
                                     coef.ctypes.data)
        return ptrow, indcol[:nblk].copy(), coef[:16 * nblk].copy()

    def spmv_bcsr4(self, ptrow, indcol, coef, x):
        nb = len(ptrow) - 1
        y = np.empty(4 * nb)
        self.l.oracle_spmv_bcsr4_fma(nb, _i32(ptrow), _i32(indcol), _f64(coef), _f64(x), y)
        return y

    def spmm_baij4(self, ptrow, indcol, coef, X, literal=False):
        """MatMatMult_SeqBAIJ_4_AVX2 restated (literal: with the reference's factor 4 from its lane reduction)."""
        nb = len(ptrow) - 1
        X = np.asfortranarray(X, dtype=np.float64)
        Y = np.zeros((4 * nb, X.shape[1]), order="F")
        self.l.oracle_spmm_baij4(nb, _i32(ptrow), _i32(indcol), _f64(coef), X.shape[1], X.ctypes.data, 4 * nb,
                                 Y.ctypes.data, 4 * nb, 1 if literal else 0)
        return Y

    # -- vectors ---------------------------------------------------------------------------
    def norm2(self, x):
        return float(self.l.oracle_norm2(len(x), _f64(x)))

    def rel_error(self, ref, test):
        return float(self.l.oracle_rel_error(len(ref), _f64(ref), _f64(test)))

    def dot(self, a, b):
        return float(self.l.oracle_dot(len(a), _f64(a), _f64(b)))

    def orthogonalize(self, x, y, alpha=1e-8):
        y = _f64(y).copy()
        beta = self.l.oracle_orthogonalize(len(x), _f64(x), y, alpha)
        return y, float(beta)

    def orthonormalize_against_basis(self, basis, y):
        """mpk/2SpMV.cpp:13-28: returns (y after the Gram-Schmidt sweep, its norm -- computed and dropped by the reference)."""
        B = np.ascontiguousarray(np.stack([_f64(b) for b in basis])) if len(basis) else np.zeros((0, len(y)))
        y = _f64(y).copy()
        nrm = self.l.oracle_orthonormalize_against_basis(len(y), len(basis), B.reshape(-1) if B.size else np.zeros(1), y)
        return y, float(nrm)

    # -- CG (parity unpinned) --------------------------------------------------------------
    def cg(self, ptrow, indcol, coef, b, tol=1e-8, maxit=1000):
        n = len(ptrow) - 1
        x = np.zeros(n)
        hist = np.zeros(maxit + 1)
        rel = C.c_double(0.0)
        it = self.l.oracle_cg(n, _i32(ptrow), _i32(indcol), _f64(coef), _f64(b), x, tol, maxit,
                              C.byref(rel), hist.ctypes.data)
        return x, it, rel.value, hist[:it + 1].copy()

    def scg(self, ptrow, indcol, coef, b, s=4, tol=1e-8, maxit=1000):
        """s-step (CA-)CG, monomial basis: CPU restatement of navierstokes_b200/csrc/sstep_cg.cu (the reference has
        no CG, SURVEY.md F2 -- parity unpinned).  Products use the pinned oracle SpMV (SpMV_CSR_FMA,
        reference mpk/SpMV.cpp:41-56); Gram and updates in numpy.  Returns (x, iterations, relres, converged)."""
        n = len(ptrow) - 1
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(n)
        r = b.copy()
        p = b.copy()
        bb = float(b @ b)
        if bb == 0.0:
            return x, 0, 0.0, True
        m = 2 * s + 1
        it = 0
        rr = bb
        done = False
        while it < maxit and not done:
            V = np.empty((m, n))
            V[0] = p
            for j in range(s):
                V[j + 1] = self.spmv(ptrow, indcol, coef, V[j])
            V[s + 1] = r
            for j in range(s - 1):
                V[s + 2 + j] = self.spmv(ptrow, indcol, coef, V[s + 1 + j])
            G = V @ V.T
            xc = np.zeros(m); rc = np.zeros(m); pc = np.zeros(m)
            pc[0] = 1.0
            rc[s + 1] = 1.0
            rr = G[s + 1, s + 1]
            for _ in range(s):
                if it >= maxit:
                    break
                w = np.zeros(m)
                w[1:s + 1] = pc[0:s]
                w[s + 2:2 * s + 1] = pc[s + 1:2 * s]
                denom = pc @ (G @ w)
                if not (denom > 0.0 and rr > 0.0):
                    done = True  # breakdown
                    break
                alpha = rr / denom
                xc += alpha * pc
                rc -= alpha * w
                rr_new = rc @ (G @ rc)
                it += 1
                if rr_new <= tol * tol * bb:
                    rr = max(rr_new, 0.0)
                    done = True
                    break
                pc = rc + (rr_new / rr) * pc
                rr = rr_new
            x = x + xc @ V
            r = rc @ V
            p = pc @ V
        return x, it, float(np.sqrt(max(rr, 0.0) / bb)), bool(rr <= tol * tol * bb)

    def true_relres(self, ptrow, indcol, coef, b, x):
        n = len(ptrow) - 1
        return float(self.l.oracle_true_relres(n, _i32(ptrow), _i32(indcol), _f64(coef), _f64(b), _f64(x)))


class _Reference:
    """Binding of oracle/_ref/libnsref_*.so (the unmodified reference, compiled)."""

    def __init__(self):
        self._s = None
        self._m = None

    def available(self) -> bool:
        return (_HERE / "_ref" / "libnsref_spmv.so").exists() and (_HERE / "_ref" / "libnsref_multi0.so").exists()

    @property
    def s(self):
        if self._s is None:
            s = _load(_HERE / "_ref" / "libnsref_spmv.so")
            s.ref_csr_new.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p]
            s.ref_csr_new.restype = C.c_void_p
            s.ref_csr_free.argtypes = [C.c_void_p]
            s.ref_spmv_csr.argtypes = [C.c_void_p, C.c_int, _f64p, _f64p]
            s.ref_spmv_csr.restype = C.c_int
            s.ref_generate_1st_layer.argtypes = [C.c_void_p, _i32p]
            s.ref_spm2v_csr.argtypes = [C.c_void_p, C.c_int, _i32p, _f64p, _f64p, _f64p]
            s.ref_spm2v_csr.restype = C.c_int
            s.ref_coo2csr.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
            s.ref_coo2csr.restype = C.c_int
            s.ref_generate_bcsr4.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_void_p, C.c_void_p, C.c_void_p]
            s.ref_generate_bcsr4.restype = C.c_int
            s.ref_spmv_bcsr4.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p]
            s.ref_spmv_bcsr4.restype = C.c_int
            s.ref_spm2v_bcsr4.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p, _f64p]
            s.ref_spm2v_bcsr4.restype = C.c_int
            s.ref_norm2.argtypes = [C.c_int, _f64p]
            s.ref_norm2.restype = C.c_double
            s.ref_rel_error.argtypes = [C.c_int, _f64p, _f64p]
            s.ref_rel_error.restype = C.c_double
            s.ref_orthogonalize.argtypes = [C.c_int, _f64p, _f64p, C.c_double]
            s.ref_orthonormalize_against_basis.argtypes = [C.c_int, C.c_int, _f64p, _f64p]
            self._s = s
        return self._s

    @property
    def m(self):
        if self._m is None:
            m = _load(_HERE / "_ref" / "libnsref_multi0.so")
            m.ref0_spmv.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
            m.ref0_spmkv.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p]
            m.ref0_spmkv.restype = C.c_int
            self._m = m
        return self._m

    VARIANTS = {"x87": 0, "opt": 1, "fma": 2, "avx2": 3}

    class Csr:
        def __init__(self, outer, ptrow, indcol, coef):
            self.o = outer
            self.n = len(ptrow) - 1
            self.nnz = len(indcol)
            self.h = outer.s.ref_csr_new(self.n, self.nnz, _i32(ptrow), _i32(indcol), _f64(coef))

        def spmv(self, x, variant="fma", out=None):
            y = np.empty(self.n) if out is None else out
            rc = self.o.s.ref_spmv_csr(self.h, self.o.VARIANTS[variant], _f64(x), y)
            assert rc == 0
            return y

        def generate_1st_layer(self):
            pe = np.empty(self.nnz, dtype=np.int32)
            self.o.s.ref_generate_1st_layer(self.h, pe)
            return pe

        def spm2v(self, x, variant="opt", ptrowend1=None):
            if ptrowend1 is None:
                ptrowend1 = self.generate_1st_layer()
            y = np.empty(self.n)
            z = np.empty(self.n)
            rc = self.o.s.ref_spm2v_csr(self.h, self.o.VARIANTS[variant], _i32(ptrowend1), _f64(x), y, z)
            assert rc == 0
            return y, z

        def close(self):
            if self.h:
                self.o.s.ref_csr_free(self.h)
                self.h = None

        def __del__(self):
            try:
                self.close()
            except Exception:
                pass

    def csr(self, ptrow, indcol, coef):
        return self.Csr(self, ptrow, indcol, coef)

    def coo2csr(self, nrow, irow, jcol, val):
        nnz = len(irow)
        ptrow = np.zeros(nrow + 1, dtype=np.int32)
        indcol = np.zeros(max(nnz, 1), dtype=np.int32)
        coef = np.zeros(max(nnz, 1), dtype=np.float64)
        kept = self.s.ref_coo2csr(nrow, nnz, _i32(irow), _i32(jcol), _f64(val), ptrow, indcol, coef)
        return ptrow, indcol[:kept].copy(), coef[:kept].copy()

    def generate_bcsr4(self, nrow, irow, jcol, val):
        nnz = len(irow)
        irow, jcol, val = _i32(irow), _i32(jcol), _f64(val)
        nblk = self.s.ref_generate_bcsr4(nrow, nnz, irow, jcol, val, None, None, None)
        ptrow = np.zeros(nrow // 4 + 1, dtype=np.int32)
        indcol = np.zeros(max(nblk, 1), dtype=np.int32)
        coef = np.zeros(16 * max(nblk, 1), dtype=np.float64)
        self.s.ref_generate_bcsr4(nrow, nnz, irow, jcol, val, ptrow.ctypes.data, indcol.ctypes.data,
                                  coef.ctypes.data)
        return ptrow, indcol[:nblk].copy(), coef[:16 * nblk].copy()

    def spmv_bcsr4(self, ptrow, indcol, coef, x, variant="fma"):
        nb = len(ptrow) - 1
        y = np.empty(4 * nb)
        rc = self.s.ref_spmv_bcsr4(nb, len(indcol), _i32(ptrow), _i32(indcol), _f64(coef),
                                   self.VARIANTS[variant], _f64(x), y)
        assert rc == 0
        return y

    def spm2v_bcsr4(self, ptrow, indcol, coef, x, variant="opt"):
        nb = len(ptrow) - 1
        y = np.empty(4 * nb)
        z = np.empty(4 * nb)
        rc = self.s.ref_spm2v_bcsr4(nb, len(indcol), _i32(ptrow), _i32(indcol), _f64(coef),
                                    self.VARIANTS[variant], _f64(x), y, z)
        assert rc == 0
        return y, z

    def norm2(self, x):
        return float(self.s.ref_norm2(len(x), _f64(x)))

    def rel_error(self, a, b):
        return float(self.s.ref_rel_error(len(a), _f64(a), _f64(b)))

    def orthogonalize(self, x, y, alpha=1e-8):
        y = _f64(y).copy()
        self.s.ref_orthogonalize(len(x), _f64(x), y, alpha)
        return y

    def orthonormalize_against_basis(self, basis, y):
        B = np.ascontiguousarray(np.stack([_f64(b) for b in basis]))
        y = _f64(y).copy()
        self.s.ref_orthonormalize_against_basis(len(y), len(basis), B.reshape(-1), y)
        return y

    def multi0_spmv(self, ptrow, indcol, coef, x):
        n = len(ptrow) - 1
        y = np.empty(n)
        self.m.ref0_spmv(n, len(indcol), _i32(ptrow), _i32(indcol), _f64(coef), _f64(x), y)
        return y

    def multi0_spmkv(self, ptrow, indcol, coef, depth, x):
        n = len(ptrow) - 1
        out = np.zeros((4, n))
        rc = self.m.ref0_spmkv(n, len(indcol), _i32(ptrow), _i32(indcol), _f64(coef), depth, _f64(x),
                               out.reshape(-1))
        assert rc == 0
        return out[:depth]


lib = _Oracle()
ref = _Reference()
