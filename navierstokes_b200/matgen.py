"""Synthetic CSR operators and vectors for the SpMV / matrix-powers / CG hot path.

The reference ships no matrices (its ``mat/matrix{1..10}_aij.mtx`` files and ``mmesh.tar.gz`` are
missing), so every workload is generated:

* ``laplace2d_5pt`` / ``laplace3d_7pt`` -- BASELINE.json configs C2, C3, C5 (diagonal 4 / 6,
  off-diagonals -1, natural lexicographic ordering, Dirichlet truncation).
* ``fem_baij4``      -- C1 substitute: a 4-dof-per-node P1-P1 block operator on a Kuhn-split tet
  mesh with the block layout of the reference's assembly (src/benchmark_spmv.c:100-118) and
  values rounded through ``float`` like the reference's Matrix-Market reader
  (mpk/SpM2V.cpp:846-850).  Every row length is a multiple of 4, as in the reference's matrices.
* ``tet_p1_laplacian`` -- C4: P1 stiffness matrix ``vol * grad(phi_i).grad(phi_j)`` (the formula of
  src/integration.c:19-57,231-236) on a Kuhn mesh, optionally node-permuted and RCM-reordered.

All index arrays are int32 (the reference's ``csrmatrix`` uses ``int``, mpk/SpMV.h:18-24), values fp64.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Csr:
    """Host CSR triple with the reference's field names (mpk/SpMV.h:18-24)."""
    n: int
    ptrow: np.ndarray   # int32[n_rows+1]
    indcol: np.ndarray  # int32[nnz]  (global column ids)
    coef: np.ndarray    # float64[nnz]
    ncols: int | None = None  # number of columns (== n for square operators)
    row0: int = 0             # global index of local row 0 (row slabs of a distributed operator)

    @property
    def nnz(self) -> int:
        return int(self.ptrow[-1])

    @property
    def nrows(self) -> int:
        return len(self.ptrow) - 1

    def spmv_bytes(self) -> int:
        """Algorithmic bytes of one product: 12*nnz + 4(n+1) + 8n + 8n."""
        n = self.nrows
        return 12 * self.nnz + 4 * (n + 1) + 16 * n

    def mpk_bytes(self, k: int) -> int:
        """Compulsory single-pass bytes of A^1..k x: 12*nnz + 4(n+1) + 8n + 8nk."""
        n = self.nrows
        return 12 * self.nnz + 4 * (n + 1) + 8 * n + 8 * n * k


def _stencil(offsets, valid, diag, n, row0, nrows):
    """Assemble rows [row0, row0+nrows) of a constant-coefficient stencil operator.

    offsets: ascending list of column offsets; valid(r, j) -> bool mask for offset j.
    """
    r = np.arange(row0, row0 + nrows, dtype=np.int64)
    m = len(offsets)
    mask = np.empty((nrows, m), dtype=bool)
    for j in range(m):
        mask[:, j] = valid(r, j)
    counts = mask.sum(axis=1, dtype=np.int64)
    ptrow = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=ptrow[1:])
    cols = np.empty((nrows, m), dtype=np.int32)
    vals = np.empty((nrows, m), dtype=np.float64)
    for j, off in enumerate(offsets):
        cols[:, j] = (r + off).astype(np.int32)
        vals[:, j] = diag if off == 0 else -1.0
    indcol = cols[mask]
    coef = vals[mask]
    assert ptrow[-1] < 2**31, "nnz must fit the reference's int (mpk/SpMV.h:20)"
    return Csr(n=n, ptrow=ptrow.astype(np.int32), indcol=indcol, coef=coef, ncols=n, row0=row0)


def laplace2d_5pt(nx: int, ny: int | None = None, row0: int = 0, nrows: int | None = None) -> Csr:
    """2D 5-point Poisson operator on an nx x ny grid (C2: nx = ny = 4096)."""
    ny = nx if ny is None else ny
    n = nx * ny
    nrows = n - row0 if nrows is None else nrows
    offsets = [-nx, -1, 0, 1, nx]

    def valid(r, j):
        ix = r % nx
        iy = r // nx
        return [iy > 0, ix > 0, np.ones_like(r, dtype=bool), ix < nx - 1, iy < ny - 1][j]

    return _stencil(offsets, valid, 4.0, n, row0, nrows)


def laplace3d_7pt(nx: int, ny: int | None = None, nz: int | None = None, row0: int = 0,
                  nrows: int | None = None) -> Csr:
    """3D 7-point Laplacian on an nx x ny x nz grid (C3: 256^3, C5: 512^3)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    nrows = n - row0 if nrows is None else nrows
    pl = nx * ny
    offsets = [-pl, -nx, -1, 0, 1, nx, pl]

    def valid(r, j):
        ix = r % nx
        iy = (r // nx) % ny
        iz = r // pl
        return [iz > 0, iy > 0, ix > 0, np.ones_like(r, dtype=bool), ix < nx - 1, iy < ny - 1,
                iz < nz - 1][j]

    return _stencil(offsets, valid, 6.0, n, row0, nrows)


def random_stencil3d(nx: int, ny: int, nz: int, seed: int = 0, max_points: int = 13, drop: float = 0.05) -> Csr:
    """Variable-coefficient stencil operator on an nx x ny x nz grid with a RANDOM stencil shape: up to `max_points`
    offsets (dx, dy, dz) in [-2, 2]^3 drawn from the seed, random fp64 coefficients per entry, neighbours outside the
    grid truncated, and a fraction `drop` of the remaining entries removed at random (ragged rows, a few empty ones).
    Structured enough that tiles reference x in a few runs, irregular enough to exercise the packing."""
    rng = np.random.default_rng(seed)
    cand = [(dx, dy, dz) for dz in (-2, -1, 0, 1, 2) for dy in (-2, -1, 0, 1, 2) for dx in (-2, -1, 0, 1, 2)]
    pick = rng.choice(len(cand), size=min(max_points, len(cand)), replace=False)
    offs = sorted({cand[i] for i in pick} | {(0, 0, 0)}, key=lambda o: (o[2], o[1], o[0]))
    n = nx * ny * nz
    r = np.arange(n, dtype=np.int64)
    ix, iy, iz = r % nx, (r // nx) % ny, r // (nx * ny)
    m = len(offs)
    mask = np.empty((n, m), dtype=bool)
    cols = np.empty((n, m), dtype=np.int64)
    for j, (dx, dy, dz) in enumerate(offs):
        ok = (ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) & (iz + dz >= 0) & (iz + dz < nz)
        mask[:, j] = ok & (rng.random(n) >= drop)
        cols[:, j] = r + dx + dy * nx + dz * nx * ny
    counts = mask.sum(axis=1, dtype=np.int64)
    ptrow = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=ptrow[1:])
    vals = rng.uniform(-2.0, 2.0, size=(n, m))
    return Csr(n=n, ptrow=ptrow.astype(np.int32), indcol=cols[mask].astype(np.int32), coef=vals[mask], ncols=n)


# ---------------------------------------------------------------------------------------------
# Tetrahedral meshes (Kuhn 6-tet split of a structured cube)
# ---------------------------------------------------------------------------------------------
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]],
                 dtype=np.int64)


def kuhn_mesh(m: int, jitter: float = 0.0, seed: int = 1):
    """(coords[(m+1)^3, 3], tets[6 m^3, 4]) for the unit cube split into m^3 cells x 6 tets."""
    g = m + 1
    idx = np.arange(g)
    X, Y, Z = np.meshgrid(idx, idx, idx, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float64) / m
    if jitter:
        rng = np.random.default_rng(seed)
        interior = np.all((coords > 0) & (coords < 1), axis=1)
        coords[interior] += rng.uniform(-jitter / m, jitter / m, size=(int(interior.sum()), 3))
    c = np.arange(m)
    I, J, K = np.meshgrid(c, c, c, indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()

    def nid(i, j, k):
        return (i * g + j) * g + k

    corners = np.stack([nid(I + ((b >> 2) & 1), J + ((b >> 1) & 1), K + (b & 1)) for b in range(8)], axis=1)
    tets = corners[:, _KUHN].reshape(-1, 4)
    return coords, tets


def _tet_gradients(coords, tets):
    """volume[ne], grad[ne, 4, 3] of the P1 basis functions (src/integration.c:7-67)."""
    p = coords[tets]                       # ne,4,3
    d = p[:, 1:, :] - p[:, :1, :]          # ne,3,3 edge vectors from vertex 0
    det = np.linalg.det(d)
    vol = np.abs(det) / 6.0
    inv = np.linalg.inv(d)                 # rows of inv^T are gradients of phi_1..3
    g123 = np.transpose(inv, (0, 2, 1))
    g0 = -g123.sum(axis=1, keepdims=True)
    return vol, np.concatenate([g0, g123], axis=1)


def _assemble(n, rows, cols, vals):
    import scipy.sparse as sp
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _from_scipy(A) -> Csr:
    assert A.nnz < 2**31
    return Csr(n=A.shape[0], ptrow=A.indptr.astype(np.int32), indcol=A.indices.astype(np.int32),
               coef=A.data.astype(np.float64), ncols=A.shape[1])


def tet_p1_laplacian(m: int, permute_seed: int | None = None, rcm: bool = False, jitter: float = 0.2) -> Csr:
    """P1 stiffness matrix on the Kuhn mesh of an m^3-cell cube (C4 uses m = 367 -> 368^3 nodes).

    A small mass-like shift (1e-3 * diag) keeps the operator non-singular (pure Neumann otherwise).
    permute_seed: random node permutation (seed 2); rcm: reverse Cuthill-McKee after it.
    """
    coords, tets = kuhn_mesh(m, jitter=jitter)
    n = coords.shape[0]
    vol, g = _tet_gradients(coords, tets)
    ke = vol[:, None, None] * np.einsum("eid,ejd->eij", g, g)      # ne,4,4
    rows = np.repeat(tets, 4, axis=1).ravel()
    cols = np.tile(tets, (1, 4)).ravel()
    A = _assemble(n, rows, cols, ke.ravel())
    import scipy.sparse as sp
    A = (A + 1e-3 * sp.diags(A.diagonal())).tocsr()
    if permute_seed is not None:
        perm = np.random.default_rng(permute_seed).permutation(n)
        A = A[perm][:, perm].tocsr()
    A.sort_indices()
    out = _from_scipy(A)
    if rcm:
        out = rcm_reorder(out)
    return out


def rcm_reorder(A: Csr) -> Csr:
    """Reverse Cuthill-McKee through the library's own host code (nsk_rcm + nsk_csr_permute, csrc/reorder.cpp)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    n = A.nrows
    ptrow = np.ascontiguousarray(A.ptrow, np.int32)
    indcol = np.ascontiguousarray(A.indcol, np.int32)
    coef = np.ascontiguousarray(A.coef, np.float64)
    perm = np.empty(n, np.int32)
    _lib.check(lib.nsk_rcm(n, ptrow.ctypes.data, indcol.ctypes.data, perm.ctypes.data))
    p2 = np.empty(n + 1, np.int32)
    c2 = np.empty(len(indcol), np.int32)
    v2 = np.empty(len(indcol), np.float64)
    _lib.check(lib.nsk_csr_permute(n, ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data, perm.ctypes.data,
                                   p2.ctypes.data, c2.ctypes.data, v2.ctypes.data))
    return Csr(n=n, ptrow=p2, indcol=c2, coef=v2, ncols=n)


def bandwidth(A: Csr) -> int:
    from . import _lib
    lib = _lib.load()
    ptrow = np.ascontiguousarray(A.ptrow, np.int32)
    indcol = np.ascontiguousarray(A.indcol, np.int32)
    return int(lib.nsk_csr_bandwidth(A.nrows, ptrow.ctypes.data, indcol.ctypes.data))


def fem_baij4(m: int, jitter: float = 0.2, float_round: bool = True) -> Csr:
    """4-dof-per-node stabilised P1-P1 operator on the Kuhn mesh of an m^3-cell cube (C1 substitute).

    Node block (src/benchmark_spmv.c:100-118): [[K+M, 0, 0, Bx^T], [0, K+M, 0, By^T], [0, 0, K+M, Bz^T],
    [-Bx, -By, -Bz, D]] with K = nu*vol*grad.grad, M lumped-consistent mass, B = vol/4 * grad,
    D = delta*h^2*vol*grad.grad (src/integration.c:84-109,212-238; Re = 100, delta = 0.05).
    All 16 entries of every node-pair block are stored (PETSc BAIJ stores full blocks, explicit zeros
    included), so every row length is a multiple of 4 -- the property the reference's AVX2 CSR kernels
    rely on.  Values are rounded through float32 when float_round.
    m = 31 gives 32768 nodes / 131072 rows, close to the reference's matrix6 (30370 nodes).
    """
    coords, tets = kuhn_mesh(m, jitter=jitter)
    nn = coords.shape[0]
    vol, g = _tet_gradients(coords, tets)
    nu, delta = 1.0 / 100.0, 0.05
    h2 = np.cbrt(6.0 * vol) ** 2
    gg = np.einsum("eid,ejd->eij", g, g)                                   # ne,4,4
    K = nu * vol[:, None, None] * gg
    Mm = vol[:, None, None] / 20.0 * (np.ones((4, 4)) + np.eye(4))[None]
    D = (delta * h2 * vol)[:, None, None] * gg
    B = vol[:, None, None] / 4.0 * g                                       # ne,4(j),3  -> (q_i, div phi_j)
    blk = np.zeros((tets.shape[0], 4, 4, 4, 4))                            # e, i, j, a, b
    for a in range(3):
        blk[:, :, :, a, a] = K + Mm
        blk[:, :, :, a, 3] = -np.broadcast_to(B[:, :, None, a], K.shape)   # grad p in momentum
        blk[:, :, :, 3, a] = np.broadcast_to(B[:, None, :, a], K.shape)    # div u in continuity
    blk[:, :, :, 3, 3] = D
    ti = tets[:, :, None, None, None]
    tj = tets[:, None, :, None, None]
    aa = np.arange(4)[None, None, None, :, None]
    bb = np.arange(4)[None, None, None, None, :]
    rows = np.broadcast_to(4 * ti + aa, blk.shape).ravel()
    cols = np.broadcast_to(4 * tj + bb, blk.shape).ravel()
    # keep explicit zeros of the block pattern: assemble pattern and values separately
    import scipy.sparse as sp
    n = 4 * nn
    V = sp.coo_matrix((blk.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    P = sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(n, n)).tocsr()
    V.sum_duplicates(); P.sum_duplicates(); V.sort_indices(); P.sort_indices()
    # V may have dropped nothing (coo->csr keeps explicit zeros), but guard by re-indexing on P's pattern
    assert V.nnz == P.nnz and np.array_equal(V.indices, P.indices)
    coef = V.data.astype(np.float32).astype(np.float64) if float_round else V.data.astype(np.float64)
    out = Csr(n=n, ptrow=V.indptr.astype(np.int32), indcol=V.indices.astype(np.int32), coef=coef, ncols=n)
    assert np.all(np.diff(out.ptrow) % 4 == 0)
    return out


def random_csr(n: int, mean_row: float, seed: int = 0, empty_rows: bool = True, max_row: int | None = None) -> Csr:
    """Ragged random operator: Poisson-distributed row lengths (some empty), sorted unique columns."""
    rng = np.random.default_rng(seed)
    lens = rng.poisson(mean_row, size=n)
    if max_row is not None:
        lens = np.minimum(lens, max_row)
    lens = np.minimum(lens, n)
    if not empty_rows:
        lens = np.maximum(lens, 1)
    ptrow = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=ptrow[1:])
    indcol = np.empty(int(ptrow[-1]), dtype=np.int32)
    for i in range(n):
        if lens[i]:
            indcol[ptrow[i]:ptrow[i + 1]] = np.sort(rng.choice(n, size=lens[i], replace=False))
    coef = rng.uniform(-1.0, 1.0, size=int(ptrow[-1]))
    return Csr(n=n, ptrow=ptrow.astype(np.int32), indcol=indcol, coef=coef, ncols=n)


def random_banded_csr(n: int, half_bw: int, mean_row: float, seed: int = 0, empty_rows: bool = True,
                      len_range: tuple[int, int] | None = None) -> Csr:
    """Ragged random operator whose columns stay within `half_bw` of the diagonal (what RCM produces):
    Poisson row lengths (some empty) or, with len_range=(lo, hi), uniformly drawn lengths with 2 % empty rows;
    sorted unique columns.  The fused matrix-powers kernels apply."""
    rng = np.random.default_rng(seed)
    if len_range is not None:
        lens = rng.integers(len_range[0], len_range[1] + 1, size=n)
        if empty_rows:
            lens[rng.random(n) < 0.02] = 0
    else:
        lens = rng.poisson(mean_row, size=n)
    lens = np.minimum(lens, half_bw + 1)  # the clipped window at a boundary still holds a row
    if not empty_rows:
        lens = np.maximum(lens, 1)
    ptrow = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=ptrow[1:])
    indcol = np.empty(int(ptrow[-1]), dtype=np.int32)
    for i in range(n):
        if lens[i]:
            lo, hi = max(0, i - half_bw), min(n, i + half_bw + 1)
            indcol[ptrow[i]:ptrow[i + 1]] = np.sort(rng.choice(hi - lo, size=int(lens[i]), replace=False)) + lo
    coef = rng.uniform(-1.0, 1.0, size=int(ptrow[-1]))
    return Csr(n=n, ptrow=ptrow.astype(np.int32), indcol=indcol, coef=coef, ncols=n)


@dataclass
class Bcsr4:
    """4x4 block CSR with row-major blocks: the layout of the reference's bcsr4x4_matrix (mpk/SpMV.h:26-33)."""
    nbrows: int
    ptrow: np.ndarray   # int32 [nbrows + 1]
    indcol: np.ndarray  # int32 [nblocks], block columns in first-appearance order
    coef: np.ndarray    # float64 [16 * nblocks]


def csr_to_bcsr4(A: Csr) -> Bcsr4:
    """Blocks a CSR operator (n a multiple of 4) the way generate_BCSR4 does (reference mpk/utils.cpp:45-95): block
    columns of a block row in order of first appearance in the row-major traversal, explicit zeros inside blocks."""
    assert A.n % 4 == 0
    nb = A.n // 4
    rows = np.repeat(np.arange(A.n, dtype=np.int64), np.diff(A.ptrow))
    cols = A.indcol.astype(np.int64)
    key = (rows // 4) * nb + cols // 4
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")       # block id in storage order -> index into uniq
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))            # uniq index -> block id
    blk = rank[inv]
    coef = np.zeros(16 * len(uniq))
    coef[blk * 16 + (rows % 4) * 4 + cols % 4] = A.coef
    brow = uniq[order] // nb
    ptrow = np.zeros(nb + 1, dtype=np.int64)
    np.add.at(ptrow, brow + 1, 1)
    ptrow = np.cumsum(ptrow)
    return Bcsr4(nbrows=nb, ptrow=ptrow.astype(np.int32), indcol=(uniq[order] % nb).astype(np.int32), coef=coef)


# ---------------------------------------------------------------------------------------------
# Input vectors
# ---------------------------------------------------------------------------------------------
def vec_ones(n: int) -> np.ndarray:
    """x = 1, the reference's practice (mpk/2SpMV.cpp:122)."""
    return np.ones(n)


def vec_sin(n: int, shift: float = 0.0, row0: int = 0) -> np.ndarray:
    """x[j] = sin(0.001 j + shift), the reference's fake Krylov vectors (mpk/2SpMV.cpp:114)."""
    return np.sin(0.001 * np.arange(row0, row0 + n, dtype=np.float64) + shift)


def vec_uniform(n: int, seed: int = 1) -> np.ndarray:
    """uniform(-1, 1), seeded."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, size=n)
