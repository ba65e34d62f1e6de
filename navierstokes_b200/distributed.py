"""Host layer of the row-partitioned (multi-GPU) operator: one process per GPU.

Planning (ghost rings, local numbering, exchange lists) is done by the C++ library (nsk_plan_*,
host-only); this module feeds it matrix rows, ships the request lists between ranks with
``torch.distributed`` (setup time only -- gloo or nccl, any backend), bootstraps the library's NCCL
communicator and wraps the distributed calls.  torch is plumbing here: rendezvous and object exchange.

The reference is single-process; nothing here replaces a reference file.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, matgen
from .api import DEVICE, EXACT_FMA, HOST, Context, DeviceVector, _ptr


# ---------------------------------------------------------------------------------------------------
# row providers: hand the planner the matrix rows of arbitrary global row ids
# ---------------------------------------------------------------------------------------------------
class GlobalCsrProvider:
    """Every rank can see the whole operator on the host (tests, small problems)."""

    def __init__(self, A: matgen.Csr):
        self.A = A
        self.n = A.n

    def rows(self, gids: np.ndarray):
        A = self.A
        lens = (A.ptrow[gids + 1] - A.ptrow[gids]).astype(np.int64)
        ptr = np.zeros(len(gids) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        idx = np.repeat(A.ptrow[gids].astype(np.int64) - ptr[:-1], lens) + np.arange(int(ptr[-1]), dtype=np.int64)
        return ptr.astype(np.int32), A.indcol[idx], A.coef[idx]


class StencilProvider:
    """Rows of the 7-point (or 5-point when nz == 1 is not used) Laplacian generated analytically."""

    def __init__(self, nx: int, ny: int, nz: int):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.n = nx * ny * nz

    def rows(self, gids: np.ndarray):
        if len(gids) == 0:
            return np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0)
        # ascending ids -> a few contiguous runs (an owned slab, or the planes above / below it)
        breaks = np.flatnonzero(np.diff(gids) != 1) + 1
        starts = np.concatenate([[0], breaks])
        ends = np.concatenate([breaks, [len(gids)]])
        ptrs, cols, vals, base = [np.zeros(1, np.int64)], [], [], 0
        for s, e in zip(starts, ends):
            part = matgen.laplace3d_7pt(self.nx, self.ny, self.nz, row0=int(gids[s]), nrows=int(e - s))
            ptrs.append(part.ptrow[1:].astype(np.int64) + base)
            base += part.nnz
            cols.append(part.indcol)
            vals.append(part.coef)
        return np.concatenate(ptrs).astype(np.int32), np.concatenate(cols), np.concatenate(vals)


# ---------------------------------------------------------------------------------------------------
# plan
# ---------------------------------------------------------------------------------------------------
class Plan:
    """nsk_plan_* wrapper.  ``Plan.build`` runs the frontier loop; ``exchange_requests`` wires peers."""

    def __init__(self, nranks: int, rank: int, row_starts, depth: int):
        self.lib = _lib.load()
        self.nranks, self.rank, self.depth = nranks, rank, depth
        self.row_starts = np.ascontiguousarray(row_starts, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(self.lib.nsk_plan_create(nranks, rank, C.c_void_p(_ptr(self.row_starts)), depth, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsk_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def build(cls, nranks, rank, row_starts, depth, provider) -> "Plan":
        p = cls(nranks, rank, row_starts, depth)
        for _ in range(depth):  # exactly `depth` rounds (owned rows, ring 1 .. ring depth-1); a ring may be empty
            cnt, rows = C.c_int(), C.c_void_p()
            _lib.check(p.lib.nsk_plan_frontier(p.h, C.byref(cnt), C.byref(rows)))
            gids = np.ctypeslib.as_array(C.cast(rows, C.POINTER(C.c_int)), shape=(cnt.value,)).copy() if cnt.value \
                else np.zeros(0, np.int32)
            ptr, cols, vals = provider.rows(gids)
            ptr = np.ascontiguousarray(ptr, np.int32)
            cols = np.ascontiguousarray(cols, np.int32)
            vals = np.ascontiguousarray(vals, np.float64)
            _lib.check(p.lib.nsk_plan_add_rows(p.h, cnt.value, C.c_void_p(_ptr(ptr)), C.c_void_p(_ptr(cols)),
                                              C.c_void_p(_ptr(vals))))
        _lib.check(p.lib.nsk_plan_finalize(p.h))
        p._read_sizes()
        return p

    def _read_sizes(self):
        no, nr, nc, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        lr = np.zeros(self.depth, np.int32)
        rs = np.zeros(self.depth + 2, np.int32)
        _lib.check(self.lib.nsk_plan_sizes(self.h, C.byref(no), C.byref(nr), C.byref(nc), C.byref(nnz),
                                           C.c_void_p(_ptr(lr)), C.c_void_p(_ptr(rs))))
        self.n_owned, self.n_rows_local, self.n_cols_local, self.nnz = no.value, nr.value, nc.value, nnz.value
        self.level_rows, self.ring_start = lr, rs

    def ghosts(self) -> np.ndarray:
        p = C.c_void_p()
        _lib.check(self.lib.nsk_plan_ghosts(self.h, C.byref(p)))
        ng = self.n_cols_local - self.n_owned
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(ng,)).copy() if ng else np.zeros(0, np.int32)

    def local_csr(self) -> matgen.Csr:
        pp, pc, pv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.nsk_plan_local_csr(self.h, C.byref(pp), C.byref(pc), C.byref(pv)))
        ptr = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_int)), shape=(self.n_rows_local + 1,)).copy()
        col = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_int)), shape=(self.nnz,)).copy() if self.nnz else np.zeros(0, np.int32)
        val = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_double)), shape=(self.nnz,)).copy() if self.nnz else np.zeros(0)
        return matgen.Csr(n=self.n_rows_local, ptrow=ptr, indcol=col, coef=val, ncols=self.n_cols_local)

    def requests(self, peer: int):
        cnt, g = C.c_int(), C.c_void_p()
        rc = np.zeros(self.depth, np.int32)
        _lib.check(self.lib.nsk_plan_requests(self.h, peer, C.byref(cnt), C.byref(g), C.c_void_p(_ptr(rc))))
        gids = np.ctypeslib.as_array(C.cast(g, C.POINTER(C.c_int)), shape=(cnt.value,)).copy() if cnt.value \
            else np.zeros(0, np.int32)
        return gids, rc

    def add_send(self, peer: int, gids, ring_counts):
        gids = np.ascontiguousarray(gids, np.int32)
        rc = np.ascontiguousarray(ring_counts, np.int32)
        _lib.check(self.lib.nsk_plan_add_send(self.h, peer, len(gids), C.c_void_p(_ptr(gids)), C.c_void_p(_ptr(rc))))

    def exchange_requests(self, dist=None, all_requests=None):
        """Tell every owner what we need from it.  ``dist`` = torch.distributed (any backend); for
        single-process emulation pass ``all_requests`` = list over ranks of {peer: (gids, ring_counts)}."""
        mine = {p: self.requests(p) for p in range(self.nranks) if p != self.rank}
        mine = {p: v for p, v in mine.items() if len(v[0])}
        if all_requests is None:
            gathered = [None] * self.nranks
            dist.all_gather_object(gathered, mine)
        else:
            gathered = all_requests
        self.sends = {}
        for src, reqs in enumerate(gathered):
            if src == self.rank or not reqs or self.rank not in reqs:
                continue
            gids, rc = reqs[self.rank]
            self.add_send(src, gids, rc)
            self.sends[src] = (np.asarray(gids), np.asarray(rc))
        self.my_requests = mine
        return mine


def slab_row_starts(n_planes: int, plane: int, nranks: int) -> np.ndarray:
    """Contiguous z-slabs with near-equal plane counts."""
    cuts = (np.arange(nranks + 1, dtype=np.int64) * n_planes) // nranks
    return (cuts * plane).astype(np.int32)


# ---------------------------------------------------------------------------------------------------
# distributed operator on the GPU
# ---------------------------------------------------------------------------------------------------
def init_comm(ctx: Context, dist) -> None:
    """Creates the library's NCCL communicator; the unique id travels over torch.distributed."""
    lib = ctx.lib
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(lib.nsk_comm_unique_id(buf), ctx.h)
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    idb = (C.c_ubyte * 128).from_buffer_copy(obj[0])
    _lib.check(lib.nsk_comm_init(ctx.h, world, rank, idb), ctx.h)


class DistOperator:
    """Row slab of a global operator on this rank's GPU (nsk_csr_create_dist)."""

    def __init__(self, ctx: Context, dist, row_starts, provider, halo_depth: int):
        self.ctx, self.dist = ctx, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.depth = halo_depth
        self.plan = Plan.build(self.world, self.rank, row_starts, halo_depth, provider)
        self.plan.exchange_requests(dist)
        if not getattr(ctx, "_comm_ready", False):
            init_comm(ctx, dist)
            ctx._comm_ready = True
        h = C.c_void_p()
        ctx._ck(ctx.lib.nsk_csr_create_dist(ctx.h, self.plan.h, C.byref(h)))
        self.h = h
        self.row_begin = int(row_starts[self.rank])
        self.n_owned = self.plan.n_owned
        self.n_cols_local = self.plan.n_cols_local
        self.n_rows_local = self.plan.n_rows_local
        self.nnz_local = self.plan.nnz
        lc = self.plan.local_csr()
        self.nnz_owned = int(lc.ptrow[self.n_owned])
        self._local_csr = lc

    @property
    def local_csr(self) -> matgen.Csr:
        """Owned rows only, local column numbering (what the CPU baseline would multiply)."""
        lc = self._local_csr
        return matgen.Csr(n=self.n_owned, ptrow=lc.ptrow[:self.n_owned + 1].copy(), indcol=lc.indcol[:self.nnz_owned],
                          coef=lc.coef[:self.nnz_owned], ncols=self.n_cols_local)

    @property
    def spmv_bytes_owned(self) -> int:
        return 12 * self.nnz_owned + 4 * (self.n_owned + 1) + 16 * self.n_owned

    def mpk_bytes_owned(self, k: int) -> int:
        return 12 * self.nnz_owned + 4 * (self.n_owned + 1) + 8 * self.n_owned + 8 * self.n_owned * k

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.nsk_csr_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # local vectors: n_cols_local doubles, owned part first
    def new_vector(self, shared: bool = False) -> DeviceVector:
        """``shared=True`` (collective: every rank, same order): the vector is mapped into the neighbours' address
        spaces and registered, so its halo travels by the push kernel over NVLink peer memory instead of NCCL."""
        v = self.ctx.zeros(self.n_cols_local)
        if shared:
            self.enable_push()
            self.ctx.sync()
            mine = self._ipc_export(v.ptr)
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, mine)
            if self._peers:
                ptrs = (C.c_void_p * len(self._peers))()
                for i, p in enumerate(self._peers):
                    ptrs[i] = self._ipc_import(gathered[p])
                self.ctx._ck(self.ctx.lib.nsk_dist_vector_register(self.h, v.ptr, ptrs))
            self.dist.barrier()  # nobody pushes before every neighbour has registered
        return v

    def _ipc_export(self, devptr) -> bytes:
        buf = (C.c_ubyte * 80)()
        self.ctx._ck(self.ctx.lib.nsk_ipc_export(self.ctx.h, devptr, buf))
        return bytes(buf)

    def _ipc_import(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 80).from_buffer_copy(handle)
        out = C.c_void_p()
        self.ctx._ck(self.ctx.lib.nsk_ipc_import(self.ctx.h, buf, C.byref(out)))
        return out.value

    def enable_push(self) -> None:
        """One-time setup of the halo push (collective): flag blocks and receive layouts are exchanged through
        torch.distributed, every neighbour's flag block is mapped here (CUDA IPC; one node only)."""
        if getattr(self, "_push_ready", False):
            return
        lib, ctx = self.ctx.lib, self.ctx
        self._peers = [lib.nsk_dist_peer_rank(self.h, i) for i in range(lib.nsk_dist_peer_count(self.h))]
        flags = C.c_void_p()
        ctx._ck(lib.nsk_dist_push_flags(self.h, C.byref(flags)))
        ctx.sync()
        layouts = {}
        for p in self._peers:
            rs, rc = np.zeros(self.depth, np.int32), np.zeros(self.depth, np.int32)
            ctx._ck(lib.nsk_dist_recv_layout(self.h, p, C.c_void_p(_ptr(rs)), C.c_void_p(_ptr(rc))))
            layouts[p] = (rs, rc)
        mine = {"flags": self._ipc_export(flags), "peers": self._peers, "layouts": layouts}
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, mine)
        for p in self._peers:
            theirs = gathered[p]
            rs, rc = theirs["layouts"][self.rank]
            rs, rc = np.ascontiguousarray(rs, np.int32), np.ascontiguousarray(rc, np.int32)
            fp = self._ipc_import(theirs["flags"])
            ctx._ck(lib.nsk_dist_push_peer(self.h, p, theirs["peers"].index(self.rank), C.c_void_p(_ptr(rs)),
                                           C.c_void_p(_ptr(rc)), C.c_void_p(fp)))
        self._push_ready = True

    def set_owned(self, v: DeviceVector, host: np.ndarray):
        host = np.ascontiguousarray(host, np.float64)
        assert host.size == self.n_owned
        self.ctx._ck(self.ctx.lib.nsk_memcpy(self.ctx.h, v.ptr, C.c_void_p(_ptr(host)), 8 * self.n_owned, 0))
        self.ctx.sync()

    def get_owned(self, v: DeviceVector) -> np.ndarray:
        out = np.empty(self.n_owned)
        self.ctx._ck(self.ctx.lib.nsk_memcpy(self.ctx.h, C.c_void_p(_ptr(out)), v.ptr, 8 * self.n_owned, 1))
        self.ctx.sync()
        return out

    def halo_exchange(self, v: DeviceVector, depth: int):
        self.ctx._ck(self.ctx.lib.nsk_halo_exchange(self.h, v.ptr, depth))

    def spmv(self, x: DeviceVector, y: DeviceVector, mode: int = EXACT_FMA):
        self.ctx._ck(self.ctx.lib.nsk_spmv(self.h, x.ptr, y.ptr, mode, DEVICE))
        return y

    def mpk(self, k: int, x: DeviceVector, levels, mode: int = EXACT_FMA):
        ptrs = (C.c_void_p * k)(*[l.ptr.value for l in levels[:k]])
        self.ctx._ck(self.ctx.lib.nsk_mpk(self.h, k, x.ptr, ptrs, mode, DEVICE))
        return levels

    def mpk_host(self, k: int, x_owned: np.ndarray, levels_owned, mode: int = EXACT_FMA):
        ptrs = (C.c_void_p * k)(*[_ptr(l) for l in levels_owned[:k]])
        self.ctx._ck(self.ctx.lib.nsk_mpk(self.h, k, C.c_void_p(_ptr(x_owned)), ptrs, mode, HOST))
        return levels_owned

    def cg(self, b_owned: np.ndarray, tol=1e-8, maxit=1000, sstep=1):
        b = np.ascontiguousarray(b_owned, np.float64)
        x = np.empty(self.n_owned)
        it, rel = C.c_int(), C.c_double()
        s = self.ctx._ck(self.ctx.lib.nsk_cg(self.h, C.c_void_p(_ptr(b)), C.c_void_p(_ptr(x)), tol, maxit, sstep,
                                             C.byref(it), C.byref(rel), HOST))
        return x, it.value, rel.value, s == 0


class DistStencil3D(DistOperator):
    """7-point Laplacian on nx x ny x nz split into z-slabs, rows generated analytically per rank."""

    def __init__(self, ctx: Context, dist, nx: int, ny: int, nz: int, halo_depth: int):
        world = dist.get_world_size()
        row_starts = slab_row_starts(nz, nx * ny, world)
        super().__init__(ctx, dist, row_starts, StencilProvider(nx, ny, nz), halo_depth)
        self.shape = (nx, ny, nz)
