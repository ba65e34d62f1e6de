"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py

Every rank builds its slab of a small 3D Laplacian, runs SpMV / matrix powers (both strategies) / CG through
the distributed C ABI and compares its owned part bit-for-bit with the CPU oracle on the global problem.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import navierstokes_b200 as nsk
    from navierstokes_b200 import matgen, distributed as nd
    import oracle
    lib = oracle.lib
    ctx = nsk.Context(local)
    odd = "--odd" in sys.argv  # odd plane size: odd local vector lengths on every rank
    nx, ny, nz, K = (47, 39, 24 * world + 1, 4) if odd else (48, 40, 24 * world, 4)
    A = matgen.laplace3d_7pt(nx, ny, nz)
    x = matgen.vec_uniform(A.n, seed=21)
    ref = lib.mpk(A.ptrow, A.indcol, A.coef, K, x)
    op = nd.DistStencil3D(ctx, dist, nx, ny, nz, halo_depth=K)
    lo, hi = op.row_begin, op.row_begin + op.n_owned
    dx = op.new_vector()
    op.set_owned(dx, x[lo:hi])
    bad = 0
    y = op.new_vector()
    op.spmv(dx, y)
    bad += not np.array_equal(op.get_owned(y).view(np.int64), ref[0][lo:hi].view(np.int64))
    used = {}
    for strat in (1, 4, 5, 0):
        ctx.set_option("mpk_kernel", strat)
        for k in (1, 2, 3, 4):
            lv = [op.new_vector() for _ in range(k)]
            op.mpk(k, dx, lv)
            used[(strat, k)] = ctx.query("last_mpk_strategy")
            for l in range(k):
                ok = np.array_equal(op.get_owned(lv[l]).view(np.int64), ref[l][lo:hi].view(np.int64))
                bad += not ok
                if not ok:
                    print(f"rank {rank}: MISMATCH strategy={strat} k={k} level={l}", flush=True)
    # host-pointer path
    outs = [np.empty(op.n_owned) for _ in range(K)]
    op.mpk_host(K, np.ascontiguousarray(x[lo:hi]), outs)
    for l in range(K):
        bad += not np.array_equal(outs[l].view(np.int64), ref[l][lo:hi].view(np.int64))
    ctx.set_option("mpk_kernel", 0)
    if rank == 0:
        print("strategy actually run per (requested, k):", used, flush=True)
    # CG
    b = lib.spmv(A.ptrow, A.indcol, A.coef, x)
    xs, it, rel, ok = op.cg(np.ascontiguousarray(b[lo:hi]), tol=1e-9, maxit=800)
    x_ref, it_ref, _, _ = lib.cg(A.ptrow, A.indcol, A.coef, b, tol=1e-9, maxit=800)
    err = float(np.max(np.abs(xs - x_ref[lo:hi])))
    cg_ok = ok and abs(it - it_ref) <= 2 and err < 1e-6 and rel <= 1e-9
    bad += not cg_ok
    xs4, it4, rel4, ok4 = op.cg(np.ascontiguousarray(b[lo:hi]), tol=1e-9, maxit=800, sstep=4)
    err4 = float(np.max(np.abs(xs4 - x_ref[lo:hi])))
    scg_ok = ok4 and abs(it4 - it_ref) <= 2 and err4 < 1e-6
    if not scg_ok:
        print(f"rank {rank}: s-step CG FAILED it={it4} rel={rel4} err={err4}", flush=True)
    bad += not scg_ok
    # ---- halo push over NVLink peer memory: registered vectors, repeated calls (epochs / acknowledgements), mixed depths
    sx = op.new_vector(shared=True)
    op.set_owned(sx, x[lo:hi])
    l0 = ctx.launch_count
    for rep in range(3):
        for k in (4, 1, 2, 4, 3):
            lv = [op.new_vector() for _ in range(k)]
            op.mpk(k, sx, lv)
            for l in range(k):
                ok = np.array_equal(op.get_owned(lv[l]).view(np.int64), ref[l][lo:hi].view(np.int64))
                bad += not ok
                if not ok:
                    print(f"rank {rank}: MISMATCH (push) rep={rep} k={k} level={l}", flush=True)
        op.spmv(sx, y)
        bad += not np.array_equal(op.get_owned(y).view(np.int64), ref[0][lo:hi].view(np.int64))
        # new contents in the same registered vector: the neighbours must see them
        x2 = matgen.vec_uniform(A.n, seed=30 + rep)
        op.set_owned(sx, x2[lo:hi])
        ref2 = lib.mpk(A.ptrow, A.indcol, A.coef, 2, x2)
        lv = [op.new_vector() for _ in range(2)]
        op.mpk(2, sx, lv)
        for l in range(2):
            bad += not np.array_equal(op.get_owned(lv[l]).view(np.int64), ref2[l][lo:hi].view(np.int64))
        op.set_owned(sx, x[lo:hi])
    push_launches = ctx.launch_count - l0
    # timing: depth-K exchange + ack, push vs NCCL (same vector, the option switches the path)
    times = {}
    lvK = [op.new_vector() for _ in range(K)]
    for name, opt in (("push", 1), ("nccl", 0)):
        ctx.set_option("halo_push", opt)
        for _ in range(5):
            op.mpk(K, sx, lvK)
        ctx.sync(); dist.barrier()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(50):
            op.mpk(K, sx, lvK)
        e1.record()
        times[name] = e0.elapsed_ms(e1) / 50 * 1e3
    ctx.set_option("halo_push", 1)
    if rank == 0:
        print(f"push path: launches={push_launches}; small-grid k={K} call: push {times['push']:.1f} us, nccl {times['nccl']:.1f} us", flush=True)

    # ---- an UNSTRUCTURED operator across the ranks: RCM'd P1 tet Laplacian, contiguous row blocks, depth-4 rings ----------
    T = matgen.tet_p1_laplacian(14, rcm=True)
    xt = matgen.vec_uniform(T.n, seed=5)
    reft = lib.mpk(T.ptrow, T.indcol, T.coef, K, xt)
    starts = (np.arange(world + 1, dtype=np.int64) * T.n // world).astype(np.int32)
    top = nd.DistOperator(ctx, dist, starts, nd.GlobalCsrProvider(T), K)
    tlo, thi = int(starts[rank]), int(starts[rank + 1])
    for shared in (False, True):
        tx = top.new_vector(shared=shared)
        top.set_owned(tx, xt[tlo:thi])
        for strat in (1, 5, 0):
            ctx.set_option("mpk_kernel", strat)
            for k in (1, 2, 4):
                lv = [top.new_vector() for _ in range(k)]
                top.mpk(k, tx, lv)
                for l in range(k):
                    ok = np.array_equal(top.get_owned(lv[l]).view(np.int64), reft[l][tlo:thi].view(np.int64))
                    bad += not ok
                    if not ok:
                        print(f"rank {rank}: MISMATCH tet mesh shared={shared} strategy={strat} k={k} level={l}", flush=True)
    ctx.set_option("mpk_kernel", 0)
    if rank == 0:
        print(f"unstructured operator (tet P1, n={T.n}) over {world} ranks: peers of rank 0 = {top.ctx.lib.nsk_dist_peer_count(top.h)}", flush=True)

    # ---- reductions: collective by default once a communicator is attached, rank-local on request -------------------
    xo = np.ascontiguousarray(x[lo:hi])
    glob = ctx.dot(xo, xo)                      # every rank calls: the sum over all ranks
    bad += not (abs(glob - float(x @ x)) <= 1e-9 * float(x @ x))
    ctx.set_option("local_reductions", 1)
    if rank == world - 1:                       # ONE rank alone: must not enter a collective
        loc = ctx.dot(xo, xo)
        bad += not (abs(loc - float(xo @ xo)) <= 1e-9 * float(xo @ xo))
    ctx.set_option("local_reductions", 0)

    t = torch.tensor([bad], device=f"cuda:{local}")
    dist.all_reduce(t)
    if rank == 0:
        print(f"dist_check world={world}: {'OK' if t.item() == 0 else 'FAILED'}  cg it={it} (oracle {it_ref}) relres={rel:.2e} "
              f"max|x-x_ref|={err:.2e}; s-step(4) it={it4} relres={rel4:.2e} max|x-x_ref|={err4:.2e}", flush=True)
    dist.destroy_process_group()
    return int(t.item() != 0)


if __name__ == "__main__":
    sys.exit(main())
