"""Host-side mirror of the reference's SpMV / matrix-powers interface on top of the C ABI.

Two layers:

* object layer -- ``Context``, ``DeviceVector``, ``CsrMatrix``: explicit residency, streams, events.
* reference-named layer -- ``csrmatrix``, ``SpMV_CSR*``, ``SpM2V_CSR*``, ``SpM3V``, ``SpM4V``, ``norm2``,
  ``rel_error``, ``orthogonalize``, ``flush_cache`` with the reference's argument order and in-place
  output convention (mpk/SpMV.h:37-66, mpk/SpM2V.cpp:80,137, mpk/SpMVmulti0.cpp:132,191), so a parity test
  reads like the reference's own drivers (mpk/2SpMV.cpp:127-293, mpk/SpM2V.cpp:884-984).

Every call goes through ``navierstokes_b200/lib/libnsk.so``; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import NskError, check

EXACT_FMA, EXACT_MULADD, FAST = 0, 1, 2
HOST, DEVICE = 0, 1
MAX_K = 16


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Context:
    """One GPU, one stream (nsk_ctx_create / nsk_ctx_destroy)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.nsk_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsk_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, s):
        return check(s, self.h)

    def sync(self):
        self._ck(self.lib.nsk_ctx_sync(self.h))

    def set_stream(self, cuda_stream: int | None):
        self._ck(self.lib.nsk_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def set_option(self, name: str, value: int):
        self._ck(self.lib.nsk_ctx_set_option(self.h, name.encode(), int(value)))

    def query(self, name: str) -> int:
        """Introspection: 'last_spmv_kernel', 'last_mpk_strategy', 'launches' (include/nsk.h nsk_ctx_query)."""
        v = C.c_int64()
        self._ck(self.lib.nsk_ctx_query(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.nsk_ctx_launch_count(self.h))

    def device_info(self) -> dict:
        sm, l2, smem, hbm = C.c_int(), C.c_int64(), C.c_int(), C.c_int64()
        self._ck(self.lib.nsk_ctx_device_info(self.h, C.byref(sm), C.byref(l2), C.byref(smem), C.byref(hbm)))
        return {"sm_count": sm.value, "l2_bytes": l2.value, "smem_optin": smem.value, "hbm_bytes": hbm.value}

    def flush_l2(self):
        self._ck(self.lib.nsk_flush_l2(self.h))

    # -- events ---------------------------------------------------------------------------------
    def event(self) -> "Event":
        return Event(self)

    # -- memory ---------------------------------------------------------------------------------
    def empty(self, n: int) -> "DeviceVector":
        return DeviceVector(self, n)

    def zeros(self, n: int) -> "DeviceVector":
        v = DeviceVector(self, n)
        self._ck(self.lib.nsk_memset0(self.h, v.ptr, 8 * n))
        return v

    def to_device(self, a: np.ndarray) -> "DeviceVector":
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = DeviceVector(self, a.size)
        v.copy_from_host(a)
        return v

    def pinned(self, n: int) -> np.ndarray:
        """float64[n] in page-locked host memory (kept alive by the returned array)."""
        p = C.c_void_p()
        self._ck(self.lib.nsk_host_alloc(self.h, 8 * max(n, 1), C.byref(p)))
        buf = (C.c_double * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=n)
        _PinnedKeeper(self, p, arr)
        return arr

    # -- vector ops -----------------------------------------------------------------------------
    def _vec(self, x):
        if isinstance(x, DeviceVector):
            return x.ptr, DEVICE, x.n
        x = np.ascontiguousarray(x, dtype=np.float64)
        return C.c_void_p(_ptr(x)), HOST, x.size

    def dot(self, a, b) -> float:
        pa, wa, n = self._vec(a)
        pb, wb, _ = self._vec(b)
        assert wa == wb
        r = C.c_double()
        self._ck(self.lib.nsk_dot(self.h, n, pa, pb, C.byref(r), wa))
        return r.value

    def norm2(self, x) -> float:
        p, w, n = self._vec(x)
        r = C.c_double()
        self._ck(self.lib.nsk_norm2(self.h, n, p, C.byref(r), w))
        return r.value

    def rel_error(self, ref, test) -> float:
        pa, wa, n = self._vec(ref)
        pb, wb, _ = self._vec(test)
        assert wa == wb
        r = C.c_double()
        self._ck(self.lib.nsk_rel_error(self.h, n, pa, pb, C.byref(r), wa))
        return r.value

    def axpy(self, a: float, x, y):
        px, wx, n = self._vec(x)
        if isinstance(y, DeviceVector):
            self._ck(self.lib.nsk_axpy(self.h, n, a, px, y.ptr, DEVICE))
        else:
            assert y.dtype == np.float64 and y.flags.c_contiguous
            self._ck(self.lib.nsk_axpy(self.h, n, a, px, C.c_void_p(_ptr(y)), HOST))

    def orthogonalize(self, x, y, alpha: float = 1e-8) -> float:
        px, wx, n = self._vec(x)
        beta = C.c_double()
        if isinstance(y, DeviceVector):
            self._ck(self.lib.nsk_orthogonalize(self.h, n, px, y.ptr, alpha, C.byref(beta), DEVICE))
        else:
            assert y.dtype == np.float64 and y.flags.c_contiguous
            self._ck(self.lib.nsk_orthogonalize(self.h, n, px, C.c_void_p(_ptr(y)), alpha, C.byref(beta), HOST))
        return beta.value


    def orthonormalize_against_basis(self, basis, y):
        """y (numpy, updated in place, or DeviceVector) swept against the basis vectors in order; returns ||y||_2."""
        m = len(basis)
        nrm = C.c_double()
        if isinstance(y, DeviceVector):
            ptrs = (C.c_void_p * max(m, 1))(*[b.ptr.value for b in basis])
            self._ck(self.lib.nsk_orthonormalize_against_basis(self.h, y.n, m, ptrs, y.ptr, C.byref(nrm), DEVICE))
            return nrm.value
        assert isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous
        bs = [np.ascontiguousarray(b, dtype=np.float64) for b in basis]
        ptrs = (C.c_void_p * max(m, 1))(*[_ptr(b) for b in bs])
        self._ck(self.lib.nsk_orthonormalize_against_basis(self.h, y.size, m, ptrs, C.c_void_p(_ptr(y)), C.byref(nrm), HOST))
        return nrm.value
    def gram(self, vectors) -> np.ndarray:
        m = len(vectors)
        ptrs = (C.c_void_p * m)()
        where = None
        keep = []
        n = None
        for i, v in enumerate(vectors):
            p, w, n = self._vec(v)
            keep.append(v)
            ptrs[i] = p.value if isinstance(p, C.c_void_p) else p
            where = w if where is None else where
            assert w == where
        G = np.zeros((m, m))
        self._ck(self.lib.nsk_gram(self.h, n, m, ptrs, G.ctypes.data_as(_lib.c_double_p), where))
        return G


class _PinnedKeeper:
    _all = {}

    def __init__(self, ctx, p, arr):
        import weakref
        self.ctx, self.p = ctx, p
        _PinnedKeeper._all[id(self)] = self
        weakref.finalize(arr, self._free)

    def _free(self):
        try:
            if self.ctx.h:
                self.ctx.lib.nsk_host_free(self.ctx.h, self.p)
        finally:
            _PinnedKeeper._all.pop(id(self), None)


class Event:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._ck(ctx.lib.nsk_event_create(ctx.h, C.byref(h)))
        self.h = h

    def record(self):
        self.ctx._ck(self.ctx.lib.nsk_event_record(self.ctx.h, self.h))
        return self

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        self.ctx._ck(self.ctx.lib.nsk_event_elapsed_ms(self.ctx.h, self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self.ctx.h and self.h:
                self.ctx.lib.nsk_event_destroy(self.ctx.h, self.h)
        except Exception:
            pass


class DeviceVector:
    """float64[n] in HBM (nsk_malloc / nsk_free)."""

    def __init__(self, ctx: Context, n: int):
        self.ctx = ctx
        self.n = int(n)
        p = C.c_void_p()
        ctx._ck(ctx.lib.nsk_malloc(ctx.h, 8 * max(self.n, 1), C.byref(p)))
        self.ptr = p

    def copy_from_host(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n
        self.ctx._ck(self.ctx.lib.nsk_memcpy(self.ctx.h, self.ptr, C.c_void_p(_ptr(a)), 8 * self.n, 0))
        self.ctx.sync()  # `a` may be a temporary

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        out = np.empty(self.n) if out is None else out
        self.ctx._ck(self.ctx.lib.nsk_memcpy(self.ctx.h, C.c_void_p(_ptr(out)), self.ptr, 8 * self.n, 1))
        self.ctx.sync()
        return out

    def free(self):
        if self.ptr and self.ctx.h:
            self.ctx.lib.nsk_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CsrMatrix:
    """Device-resident CSR operator (nsk_csr_create); mirrors ``struct csrmatrix`` (mpk/SpMV.h:18-24)."""

    def __init__(self, ctx: Context, ptrow, indcol, coef, n_cols: int | None = None):
        self.ctx = ctx
        ptrow = np.ascontiguousarray(ptrow, dtype=np.int32)
        indcol = np.ascontiguousarray(indcol, dtype=np.int32)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        self.n = len(ptrow) - 1
        self.n_cols = self.n if n_cols is None else int(n_cols)
        self.nnz = int(ptrow[-1]) if len(ptrow) else 0
        h = C.c_void_p()
        ctx._ck(ctx.lib.nsk_csr_create(ctx.h, self.n, self.n_cols, self.nnz, C.c_void_p(_ptr(ptrow)),
                                       C.c_void_p(_ptr(indcol)), C.c_void_p(_ptr(coef)), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.nsk_csr_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def spmv_bytes(self) -> int:
        return int(self.ctx.lib.nsk_csr_spmv_bytes(self.h))

    @property
    def packed_bytes(self) -> int:
        """Bytes of the tile-packed copy the default kernels stream (0: operator does not pack, CSR kernels run)."""
        return int(self.ctx.lib.nsk_csr_packed_bytes(self.h))

    @property
    def tile_bytes(self) -> int:
        """Bytes of the sliced-ELL tile copy the default kernels stream (0: not stored that way)."""
        return int(self.ctx.lib.nsk_csr_tile_bytes(self.h))

    def mpk_bytes(self, k: int) -> int:
        return int(self.ctx.lib.nsk_csr_mpk_bytes(self.h, k))

    def spmv(self, x, y=None, mode: int = EXACT_FMA):
        """y = A x.  numpy in -> numpy out (host call, copies inside); DeviceVector in -> enqueue only."""
        if isinstance(x, DeviceVector):
            y = self.ctx.empty(self.n) if y is None else y
            self.ctx._ck(self.ctx.lib.nsk_spmv(self.h, x.ptr, y.ptr, mode, DEVICE))
            return y
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n_cols
        y = np.empty(self.n) if y is None else y
        assert y.dtype == np.float64 and y.flags.c_contiguous and y.size == self.n
        self.ctx._ck(self.ctx.lib.nsk_spmv(self.h, C.c_void_p(_ptr(x)), C.c_void_p(_ptr(y)), mode, HOST))
        return y

    def mpk(self, k: int, x, levels=None, mode: int = EXACT_FMA):
        """levels[l] = A^(l+1) x, l = 0..k-1 (the reference's y, z, w, v)."""
        ptrs = (C.c_void_p * k)()
        if isinstance(x, DeviceVector):
            levels = [self.ctx.empty(self.n) for _ in range(k)] if levels is None else levels
            for l in range(k):
                ptrs[l] = levels[l].ptr.value
            self.ctx._ck(self.ctx.lib.nsk_mpk(self.h, k, x.ptr, ptrs, mode, DEVICE))
            return levels
        x = np.ascontiguousarray(x, dtype=np.float64)
        if levels is None:
            levels = np.empty((k, self.n))
        rows = [levels[l] for l in range(k)]
        for l in range(k):
            assert rows[l].dtype == np.float64 and rows[l].flags.c_contiguous and rows[l].size == self.n
            ptrs[l] = _ptr(rows[l])
        self.ctx._ck(self.ctx.lib.nsk_mpk(self.h, k, C.c_void_p(_ptr(x)), ptrs, mode, HOST))
        return levels

    def mpk_multi(self, k: int, xs, levels=None, mode: int = EXACT_FMA):
        """levels[v][l] = A^(l+1) xs[v] for several right-hand sides.  DeviceVectors: enqueue only, two vectors per
        fused launch.  numpy arrays: host call, returns an array [nvec, k, n]."""
        nvec = len(xs)
        if isinstance(xs[0], DeviceVector):
            if levels is None:
                levels = [[self.ctx.empty(self.n) for _ in range(k)] for _ in range(nvec)]
            xp = (C.c_void_p * nvec)(*[x.ptr.value for x in xs])
            lp = (C.c_void_p * (nvec * k))(*[levels[v][l].ptr.value for v in range(nvec) for l in range(k)])
            self.ctx._ck(self.ctx.lib.nsk_mpk_multi(self.h, k, nvec, xp, lp, mode, DEVICE))
            return levels
        xs = [np.ascontiguousarray(x, dtype=np.float64) for x in xs]
        out = np.empty((nvec, k, self.n))
        xp = (C.c_void_p * nvec)(*[_ptr(x) for x in xs])
        lp = (C.c_void_p * (nvec * k))(*[out[v, l].ctypes.data for v in range(nvec) for l in range(k)])
        self.ctx._ck(self.ctx.lib.nsk_mpk_multi(self.h, k, nvec, xp, lp, mode, HOST))
        return out

    def cg(self, b, x=None, tol: float = 1e-8, maxit: int = 1000, sstep: int = 1):
        """Solves A x = b; returns (x, iterations, relres, converged)."""
        it, rel = C.c_int(), C.c_double()
        if isinstance(b, DeviceVector):
            x = self.ctx.empty(self.n) if x is None else x
            s = self.ctx._ck(self.ctx.lib.nsk_cg(self.h, b.ptr, x.ptr, tol, maxit, sstep, C.byref(it), C.byref(rel),
                                                 DEVICE))
        else:
            b = np.ascontiguousarray(b, dtype=np.float64)
            x = np.empty(self.n) if x is None else x
            s = self.ctx._ck(self.ctx.lib.nsk_cg(self.h, C.c_void_p(_ptr(b)), C.c_void_p(_ptr(x)), tol, maxit, sstep,
                                                 C.byref(it), C.byref(rel), HOST))
        return x, it.value, rel.value, s == 0


class Bcsr4Matrix:
    """Device-resident 4x4 block CSR operator (row-major blocks): the reference's bcsr4x4_matrix
    (mpk/SpMV.h:26-33) as built by generate_BCSR4 (mpk/utils.cpp:45-95)."""

    def __init__(self, ctx: Context, ptrow, indcol, coef):
        self.ctx = ctx
        ptrow = np.ascontiguousarray(ptrow, dtype=np.int32)
        indcol = np.ascontiguousarray(indcol, dtype=np.int32)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        self.nbrows = len(ptrow) - 1
        self.nblocks = len(indcol)
        self.n = 4 * self.nbrows
        assert coef.size == 16 * self.nblocks
        h = C.c_void_p()
        ctx._ck(ctx.lib.nsk_bcsr4_create(ctx.h, self.nbrows, self.nblocks, C.c_void_p(_ptr(ptrow)),
                                         C.c_void_p(_ptr(indcol)), C.c_void_p(_ptr(coef)), C.byref(h)))
        self.h = h

    def close(self):
        if self.h is not None and self.ctx.h is not None:
            self.ctx.lib.nsk_bcsr4_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def spmv_bytes(self) -> int:
        """Algorithmic bytes of one product: 128 B of values + 4 B of index per block, row pointers, x and y once."""
        return 132 * self.nblocks + 4 * (self.nbrows + 1) + 16 * self.n

    def spmv(self, x, y=None, mode: int = EXACT_FMA):
        """y = B x.  numpy in -> numpy out (host call); DeviceVector in -> enqueue only."""
        if isinstance(x, DeviceVector):
            y = self.ctx.empty(self.n) if y is None else y
            self.ctx._ck(self.ctx.lib.nsk_spmv_bcsr4(self.h, x.ptr, y.ptr, mode, DEVICE))
            return y
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        y = np.empty(self.n) if y is None else y
        self.ctx._ck(self.ctx.lib.nsk_spmv_bcsr4(self.h, C.c_void_p(_ptr(x)), C.c_void_p(_ptr(y)), mode, HOST))
        return y

    def mpk(self, k: int, x, levels=None, mode: int = EXACT_FMA):
        """levels[l] = B^(l+1) x (SpM2V_BCSR* of the reference for k = 2, mpk/SpM2V.cpp:376-801).  numpy in -> list of numpy
        out; DeviceVector in -> enqueue only."""
        if isinstance(x, DeviceVector):
            levels = [self.ctx.empty(self.n) for _ in range(k)] if levels is None else levels
            arr = (C.c_void_p * k)(*[l.ptr for l in levels])
            self.ctx._ck(self.ctx.lib.nsk_bcsr4_mpk(self.h, k, x.ptr, arr, mode, DEVICE))
            return levels
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        levels = [np.empty(self.n) for _ in range(k)] if levels is None else levels
        arr = (C.c_void_p * k)(*[C.c_void_p(_ptr(l)) for l in levels])
        self.ctx._ck(self.ctx.lib.nsk_bcsr4_mpk(self.h, k, C.c_void_p(_ptr(x)), arr, mode, HOST))
        return levels

    def spmm(self, X: np.ndarray) -> np.ndarray:
        """Y = B X for the s columns of X (n x s): MatMatMult_SeqBAIJ_4_AVX2 (src/kernels/spmm_avx2.c:7-109)."""
        X = np.asfortranarray(X, dtype=np.float64)
        assert X.ndim == 2 and X.shape[0] == self.n
        s = X.shape[1]
        Y = np.zeros((self.n, s), order="F")
        self.ctx._ck(self.ctx.lib.nsk_spmm_bcsr4(self.h, s, C.c_void_p(_ptr(X)), self.n, C.c_void_p(_ptr(Y)), self.n, HOST))
        return Y

    def krylov_basis(self, v0: np.ndarray, s: int) -> np.ndarray:
        """[v0, B v0, ..., B^s v0] (n x (s+1)): BuildKrylovBasis_AVX2 (src/kernels/spmm_avx2.c:112-168)."""
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        assert v0.size == self.n
        V = np.zeros((self.n, s + 1), order="F")
        self.ctx._ck(self.ctx.lib.nsk_krylov_basis_bcsr4(self.h, s, C.c_void_p(_ptr(v0)), C.c_void_p(_ptr(V)), self.n, HOST))
        return V


# =================================================================================================
# reference-named layer
# =================================================================================================
_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        _default_ctx = Context(0)
    return _default_ctx


@dataclass
class csrmatrix:
    """Field-for-field the reference's container (mpk/SpMV.h:18-24); arrays are numpy, host-side."""
    n: int = 0
    nnz: int = 0
    ptrow: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    indcol: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    coef: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    _gpu: CsrMatrix | None = field(default=None, repr=False, compare=False)
    _key: tuple | None = field(default=None, repr=False, compare=False)

    def gpu(self) -> CsrMatrix:
        """Device copy, uploaded on first use and re-uploaded if the host arrays were replaced."""
        key = (self.ptrow.ctypes.data, self.indcol.ctypes.data, self.coef.ctypes.data, self.n, len(self.indcol))
        if self._gpu is None or self._key != key:
            self._gpu = CsrMatrix(default_context(), self.ptrow, self.indcol, self.coef)
            self._key = key
        return self._gpu


def _out(a, n):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.size >= n
    return a


def SpMV_CSR(y, x, A: csrmatrix):
    """y = A x.  The reference's x87 variant (mpk/SpMV.cpp:6-20) is not reproducible on a GPU; this is the
    separately-rounded multiply-add chain, which is what SSE2 evaluation of the same source gives."""
    A.gpu().spmv(x, _out(y, A.n), EXACT_MULADD)


def SpMV_CSR_OPT(y, x, A: csrmatrix):
    """mpk/SpMV.cpp:23-38 (compiles to an fma chain)."""
    A.gpu().spmv(x, _out(y, A.n), EXACT_FMA)


def SpMV_CSR_FMA(y, x, A: csrmatrix):
    """mpk/SpMV.cpp:41-56 -- the bit-exact oracle flavour."""
    A.gpu().spmv(x, _out(y, A.n), EXACT_FMA)


def SpMV_CSR_AVX2(y, x, A: csrmatrix):
    """mpk/SpMV.cpp:59-85 -- reassociated; mapped to the fast mode (any row length is fine here)."""
    A.gpu().spmv(x, _out(y, A.n), FAST)


@dataclass
class bcsr4x4_matrix:
    """Field-for-field the reference's block container (mpk/SpMV.h:26-33); coef is flat, 16 doubles per block,
    row-major inside a block."""
    nrows: int = 0
    nblocks: int = 0
    ptrow: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    indcol: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    coef: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    _gpu: Bcsr4Matrix | None = field(default=None, repr=False, compare=False)
    _key: tuple | None = field(default=None, repr=False, compare=False)

    def gpu(self) -> Bcsr4Matrix:
        key = (self.ptrow.ctypes.data, self.indcol.ctypes.data, self.coef.ctypes.data, self.nrows, len(self.indcol))
        if self._gpu is None or self._key != key:
            self._gpu = Bcsr4Matrix(default_context(), self.ptrow, self.indcol, self.coef)
            self._key = key
        return self._gpu


def SpMV_BCSR(y, x, A: bcsr4x4_matrix):
    """mpk/SpMV.cpp:90-121 (x87 in the reference) -> separately rounded multiply-add chain in (block, j) order."""
    A.gpu().spmv(x, _out(y, 4 * A.nrows), EXACT_MULADD)


def SpMV_BCSR_OPT(y, x, A: bcsr4x4_matrix):
    """mpk/SpMV.cpp:124-150 (compiles to fma chains)."""
    A.gpu().spmv(x, _out(y, 4 * A.nrows), EXACT_FMA)


def SpMV_BCSR_FMA(y, x, A: bcsr4x4_matrix):
    """mpk/SpMV.cpp:153-178 -- the bit-exact flavour."""
    A.gpu().spmv(x, _out(y, 4 * A.nrows), EXACT_FMA)


def SpMV_BCSR_AVX2(y, x, A: bcsr4x4_matrix):
    """mpk/SpMV.cpp:181-219: lane i of the AVX2 kernel runs the same (block, j) fma chain as row 4*bi+i here."""
    A.gpu().spmv(x, _out(y, 4 * A.nrows), EXACT_FMA)


def _spm2v_bcsr(z, y, x, A: bcsr4x4_matrix, mode):
    A.gpu().mpk(2, x, [_out(y, 4 * A.nrows), _out(z, 4 * A.nrows)], mode)


def SpM2V_BCSR(z, y, x, A: bcsr4x4_matrix, ptrowendB=None):
    """mpk/SpM2V.cpp:376 (x87 in the reference) -> multiply-add chain."""
    _spm2v_bcsr(z, y, x, A, EXACT_MULADD)


def SpM2V_BCSR_OPT(z, y, x, A: bcsr4x4_matrix, ptrowendB=None):
    """mpk/SpM2V.cpp:475 -- y = A x, z = A y on the block operator (two products; the reference's first-touch
    schedule only changes the order in which rows of y are produced, not their values)."""
    _spm2v_bcsr(z, y, x, A, EXACT_FMA)


def SpM2V_BCSR_FMA(z, y, x, A: bcsr4x4_matrix, ptrowendB=None):
    """mpk/SpM2V.cpp:567."""
    _spm2v_bcsr(z, y, x, A, EXACT_FMA)


def SpM2V_BCSR_AVX2(z, y, x, A: bcsr4x4_matrix, ptrowendB=None):
    """mpk/SpM2V.cpp:675: lane i of the AVX2 kernel runs row 4*bi+i's (block, j) fma chain."""
    _spm2v_bcsr(z, y, x, A, EXACT_FMA)


def MatMatMult_SeqBAIJ_4_AVX2(A: bcsr4x4_matrix, X, Y, s_step: int):
    """Y = A X on s_step dense columns (reference src/kernels/spmm_avx2.c:7-109; returns A X, not the reference's 4 A X)."""
    Y[:, :s_step] = A.gpu().spmm(np.asarray(X)[:, :s_step])


def BuildKrylovBasis_AVX2(A: bcsr4x4_matrix, v0, s_step: int):
    """V = [v0, A v0, ..., A^s v0] (reference src/kernels/spmm_avx2.c:112-168)."""
    return A.gpu().krylov_basis(v0, s_step)


def Generate1stlayer_BCSR4(ptrowendB, A: bcsr4x4_matrix):
    """mpk/SpM2V.cpp:28-46: not needed by the GPU kernels; kept so drivers written against the reference run."""
    return None


# ---- ingest (host only): the reference's converters and its Matrix Market reader ------------------------------
def COO2CSR(A: csrmatrix, nrow: int, irow, jcol, val):
    """mpk/utils.cpp:97-127 (+ generate_CSR :5-43): fills A from 0-based COO; columns ascending per row, a repeated
    (i,j) is dropped (first wins); A.nnz keeps the COO count like the reference (utils.cpp:100)."""
    lib = _lib.load()
    irow = np.ascontiguousarray(irow, dtype=np.int32)
    jcol = np.ascontiguousarray(jcol, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    nnz = len(irow)
    ptrow = np.zeros(nrow + 1, dtype=np.int32)
    indcol = np.zeros(max(nnz, 1), dtype=np.int32)
    coef = np.zeros(max(nnz, 1), dtype=np.float64)
    kept = lib.nsk_coo2csr(nrow, nnz, C.c_void_p(_ptr(irow)), C.c_void_p(_ptr(jcol)), C.c_void_p(_ptr(val)),
                           C.c_void_p(_ptr(ptrow)), C.c_void_p(_ptr(indcol)), C.c_void_p(_ptr(coef)))
    if kept < 0:
        raise NskError(int(kept), "nsk_coo2csr: bad COO input")
    A.n, A.nnz = nrow, nnz
    A.ptrow, A.indcol, A.coef = ptrow, indcol[:kept].copy(), coef[:kept].copy()
    return A


def generate_BCSR4(nrow: int, irow, jcol, val) -> "bcsr4x4_matrix":
    """mpk/utils.cpp:45-95: 4x4 block CSR from 0-based COO (nrow a multiple of 4): block columns in first-appearance
    order, row-major blocks, explicit zeros, a repeated (i,j) overwrites.  nblocks is set to 0 like the reference
    (utils.cpp:78); len(indcol) is the real count."""
    lib = _lib.load()
    irow = np.ascontiguousarray(irow, dtype=np.int32)
    jcol = np.ascontiguousarray(jcol, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    nnz = len(irow)
    args = (nrow, nnz, C.c_void_p(_ptr(irow)), C.c_void_p(_ptr(jcol)), C.c_void_p(_ptr(val)))
    nblk = lib.nsk_coo2bcsr4(*args, None, None, None)
    if nblk < 0:
        raise NskError(int(nblk), "nsk_coo2bcsr4: bad COO input")
    ptrow = np.zeros(nrow // 4 + 1, dtype=np.int32)
    indcol = np.zeros(max(nblk, 1), dtype=np.int32)
    coef = np.zeros(16 * max(nblk, 1), dtype=np.float64)
    lib.nsk_coo2bcsr4(*args, C.c_void_p(_ptr(ptrow)), C.c_void_p(_ptr(indcol)), C.c_void_p(_ptr(coef)))
    return bcsr4x4_matrix(nrows=nrow // 4, nblocks=0, ptrow=ptrow, indcol=indcol[:nblk].copy(), coef=coef[:16 * nblk].copy())


def read_mtx(path: str):
    """The reader every reference driver inlines (mpk/SpM2V.cpp:815-852): returns (nrow, irow, jcol, val), 0-based,
    values rounded through float32 exactly like fscanf("%f")."""
    lib = _lib.load()
    nrow, nnz = C.c_int(), C.c_int64()
    pi, pj, pv = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    st = lib.nsk_mtx_read(str(path).encode(), C.byref(nrow), C.byref(nnz), C.byref(pi), C.byref(pj), C.byref(pv))
    if st != 0:
        raise NskError(int(st), f"cannot read Matrix Market file {path}")
    m = nnz.value
    try:
        irow = np.ctypeslib.as_array(pi, shape=(max(m, 1),))[:m].copy()
        jcol = np.ctypeslib.as_array(pj, shape=(max(m, 1),))[:m].copy()
        val = np.ctypeslib.as_array(pv, shape=(max(m, 1),))[:m].copy()
    finally:
        lib.nsk_mtx_free(pi, pj, pv)
    return nrow.value, irow, jcol, val


def Generate1stlayer(ptrowend1, A: csrmatrix):
    """The reference's first-touch schedule (mpk/SpM2V.cpp:5-26) is not needed by the GPU kernels; kept as a
    no-op so drivers written against the reference run unchanged (the plan lives in nsk_csr_create)."""
    return None


def _spmkv(levels_out, x, A: csrmatrix, mode):
    k = len(levels_out)
    for l in levels_out:
        _out(l, A.n)
    A.gpu().mpk(k, x, levels_out, mode)


def SpM2V_CSR(z, y, x, A: csrmatrix, ptrowend1=None):
    """mpk/SpM2V.cpp:80 (x87) -> multiply-add chain."""
    _spmkv([y, z], x, A, EXACT_MULADD)


def SpM2V_CSR_OPT(z, y, x, A: csrmatrix, ptrowend1=None):
    """mpk/SpM2V.cpp:137: y = A x, z = A y.  Both levels as fma chains (= two SpMV_CSR_FMA)."""
    _spmkv([y, z], x, A, EXACT_FMA)


def SpM2V_CSR_AVX2(z, y, x, A: csrmatrix, ptrowend1=None):
    """mpk/SpM2V.cpp:279 -> fast mode."""
    _spmkv([y, z], x, A, FAST)


def SpM2V0(z, y, x, A: csrmatrix, ptrowend1=None):
    """mpk/SpMVmulti0.cpp:44 (no-fma in the reference) -> multiply-add chain."""
    _spmkv([y, z], x, A, EXACT_MULADD)


def SpM2V(z, y, x, A: csrmatrix, ptrowend1=None):
    """mpk/SpMVmulti0.cpp:65 (unrolled variant of SpM2V0)."""
    _spmkv([y, z], x, A, EXACT_MULADD)


def SpM3V(w, z, y, x, A: csrmatrix, ptrowend1=None, ptrowend2=None):
    """mpk/SpMVmulti0.cpp:132."""
    _spmkv([y, z, w], x, A, EXACT_FMA)


def SpM4V(v, w, z, y, x, A: csrmatrix, ptrowend1=None, ptrowend2=None, ptrowend3=None):
    """mpk/SpMVmulti0.cpp:191."""
    _spmkv([y, z, w, v], x, A, EXACT_FMA)


def norm2(x) -> float:
    """mpk/utils.cpp:131-136."""
    return default_context().norm2(x)


def rel_error(ref, test) -> float:
    """mpk/utils.cpp:138-143."""
    return default_context().rel_error(ref, test)


def orthogonalize(nrow, x, y, alpha: float = 1e-8):
    """mpk/2SpMV.cpp:3-11: y -= alpha * <x,y> * x, in place."""
    return default_context().orthogonalize(x[:nrow], y[:nrow] if y.size != nrow else y, alpha)


def orthonormalize_against_basis(nrow, basis, y):
    """mpk/2SpMV.cpp:13-28: Gram-Schmidt sweep of y (in place) against every vector of `basis` in order; returns the
    norm the reference computes and drops (y is not scaled, as in the reference)."""
    return default_context().orthonormalize_against_basis([b[:nrow] for b in basis], y[:nrow] if y.size != nrow else y)


def flush_cache():
    """mpk/utils.cpp:146-154 -> scrub the GPU's L2."""
    default_context().flush_l2()
