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
    t = torch.tensor([bad], device=f"cuda:{local}")
    dist.all_reduce(t)
    if rank == 0:
        print(f"dist_check world={world}: {'OK' if t.item() == 0 else 'FAILED'}  cg it={it} (oracle {it_ref}) relres={rel:.2e} "
              f"max|x-x_ref|={err:.2e}; s-step(4) it={it4} relres={rel4:.2e} max|x-x_ref|={err4:.2e}", flush=True)
    dist.destroy_process_group()
    return int(t.item() != 0)


if __name__ == "__main__":
    sys.exit(main())
