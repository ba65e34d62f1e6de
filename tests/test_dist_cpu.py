"""CPU, world_size 2 (gloo): the distributed plan end to end -- ghost rings, request exchange, halo
exchange by the plan's lists, redundant ghost levels -- against the global oracle."""
import os
import sys
import traceback
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, case, errq):
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from navierstokes_b200 import matgen, distributed as nd
        import dist_emulation as emu
        lib = oracle.lib
        if case == "stencil":
            nx, ny, nz, K = 7, 6, 11, 4
            A = matgen.laplace3d_7pt(nx, ny, nz)
            rs = nd.slab_row_starts(nz, nx * ny, world)
            provider = nd.StencilProvider(nx, ny, nz)
        elif case == "tet_rcm":
            A = matgen.tet_p1_laplacian(7, permute_seed=2, rcm=True)
            K = 3
            rs = np.array([0, A.n // 2 + 5, A.n], np.int32)
            provider = nd.GlobalCsrProvider(A)
        else:  # uneven split with an empty middle-free layout: rank 1 owns a single plane
            nx, ny, nz, K = 5, 5, 9, 2
            A = matgen.laplace3d_7pt(nx, ny, nz)
            rs = np.array([0, 8 * 25, 9 * 25], np.int32)
            provider = nd.GlobalCsrProvider(A)
        plan = nd.Plan.build(world, rank, rs, K, provider)
        plan.exchange_requests(dist)
        assert plan.n_owned == rs[rank + 1] - rs[rank]
        assert plan.level_rows[K - 1] == plan.n_owned and plan.level_rows[0] == plan.n_rows_local
        # ghost ids are disjoint from the owned range and ascending inside each ring
        g = plan.ghosts()
        assert not np.any((g >= rs[rank]) & (g < rs[rank + 1]))
        for r in range(1, K + 1):
            ring = g[plan.ring_start[r] - plan.n_owned: plan.ring_start[r + 1] - plan.n_owned]
            assert np.all(np.diff(ring) > 0)
        x = matgen.vec_uniform(A.n, seed=11)
        ref = lib.mpk(A.ptrow, A.indcol, A.coef, K, x)
        for k in range(1, K + 1):
            got = emu.mpk(plan, lib, x[rs[rank]:rs[rank + 1]], k)
            for l in range(k):
                want = ref[l][rs[rank]:rs[rank + 1]]
                assert np.array_equal(got[l].view(np.int64), want.view(np.int64)), (case, rank, k, l)
        # CG: distributed recurrence reaches the same solution as the global oracle
        b = lib.spmv(A.ptrow, A.indcol, A.coef, x)
        xs, it, rel = emu.cg(plan, lib, b[rs[rank]:rs[rank + 1]], 1e-10, 500)
        x_ref, it_ref, _, _ = lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-10, maxit=500)
        assert abs(it - it_ref) <= 2 and rel <= 1e-10
        assert np.max(np.abs(xs - x_ref[rs[rank]:rs[rank + 1]])) <= 1e-7
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


@pytest.mark.parametrize("case", ["stencil", "tet_rcm", "uneven"])
def test_distributed_plan_world2(case):
    import oracle
    from navierstokes_b200 import build
    oracle.build(ref=False)
    build.build()
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"stencil": 0, "tet_rcm": 1, "uneven": 2}[case]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, errq)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errs.append("timeout")
    assert not errs and all(p.exitcode == 0 for p in procs), "\n".join(errs)


def test_plan_argument_errors():
    import ctypes as C
    from navierstokes_b200 import _lib, build
    build.build()
    lib = _lib.load()
    h = C.c_void_p()
    rs = np.array([0, 10, 5], np.int32)  # decreasing
    assert lib.nsk_plan_create(2, 0, C.c_void_p(rs.ctypes.data), 2, C.byref(h)) == -1
    rs = np.array([0, 10, 20], np.int32)
    assert lib.nsk_plan_create(2, 0, C.c_void_p(rs.ctypes.data), 0, C.byref(h)) == -1   # depth < 1
    assert lib.nsk_plan_create(2, 0, C.c_void_p(rs.ctypes.data), 2, C.byref(h)) == 0
    assert lib.nsk_plan_finalize(h) == -1  # rows not supplied yet
    lib.nsk_plan_destroy(h)


def test_plan_loop_with_empty_rings():
    """The loop nsk.h documents: exactly `depth` frontier / add_rows rounds, empty rings included (a rank of a
    block-diagonal operator has no ghosts at all; stopping at the first empty frontier would leave the plan unfinalizable)."""
    from navierstokes_b200 import build, matgen, distributed as nd
    build.build()
    n = 12
    ptr = np.arange(n + 1, dtype=np.int32)
    A = matgen.Csr(n=n, ptrow=ptr, indcol=np.arange(n, dtype=np.int32), coef=np.full(n, 2.0), ncols=n)  # diagonal
    rs = np.array([0, 5, 12], np.int32)
    for rank in range(2):
        plan = nd.Plan.build(2, rank, rs, 3, nd.GlobalCsrProvider(A))
        assert plan.n_owned == rs[rank + 1] - rs[rank] == plan.n_rows_local == plan.n_cols_local
        assert len(plan.ghosts()) == 0 and list(plan.level_rows) == [plan.n_owned] * 3
        plan.close()
    # one coupling only: rank 0's last row reads rank 1's first -> ring 1 has one entry, ring 2 (depth 3) is empty on rank 0
    lens = np.ones(n, np.int64)
    lens[4] = 2
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    col = np.concatenate([np.arange(4), [4, 5], np.arange(5, 12)]).astype(np.int32)
    B = matgen.Csr(n=n, ptrow=ptr, indcol=col, coef=np.ones(len(col)), ncols=n)
    assert B.ptrow[-1] == len(col)
    plan = nd.Plan.build(2, 0, rs, 3, nd.GlobalCsrProvider(B))
    assert list(plan.ghosts()) == [5] and plan.n_rows_local == 6 and plan.n_cols_local == 6
    plan.close()
