"""Pressure-Poisson CG across GPUs (BASELINE.json config 5 shape): one rank per GPU under torchrun,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/dist_cg.py [--nx 512 --ny 512 --planes-per-gpu 64]

7-point Laplacian nx x ny x (planes-per-gpu * world) split into z-slabs (8 GPUs x 64 planes = 512^3), b = A x_true,
x0 = 0, classical CG and s-step CG (s = 4) to ||r||/||b|| <= 1e-8.  Prints iterations, time, iterations/s and the
error against x_true.  Host wall clock around nsk_cg with owned parts of b and x in host memory (copies included).
"""
import argparse
import datetime
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=512)
    ap.add_argument("--planes-per-gpu", type=int, default=64)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--maxit", type=int, default=4000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    import navierstokes_b200 as nsk
    from navierstokes_b200 import distributed as nd
    ctx = nsk.Context(local)
    nz = args.planes_per_gpu * world
    t0 = time.time()
    op = nd.DistStencil3D(ctx, dist, args.nx, args.ny, nz, halo_depth=4)
    t_plan = time.time() - t0
    gidx = op.row_begin + np.arange(op.n_owned)
    x_true = np.sin(0.001 * gidx) + 0.5
    dx, db = op.new_vector(), op.new_vector()
    op.set_owned(dx, x_true)
    op.spmv(dx, db)
    b = op.get_owned(db)
    if rank == 0:
        print(f"# {args.nx}x{args.ny}x{nz} on {world} GPUs: {op.n_owned} owned rows per rank, plan {t_plan:.1f}s", flush=True)
    for s in (1, 4):
        op.cg(b, tol=1e-300, maxit=8, sstep=s)  # warm: plans, packed operator, workspaces
        ctx.sync()
        dist.barrier()
        t0 = time.perf_counter()
        x, it, rel, ok = op.cg(b, tol=args.tol, maxit=args.maxit, sstep=s)
        ctx.sync()
        dist.barrier()
        dt = time.perf_counter() - t0
        err = torch.tensor([float(np.max(np.abs(x - x_true)))], device=f"cuda:{local}")
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            name = "classical CG" if s == 1 else f"s-step CG s={s}"
            print(f"{name:16s}: {it} iterations to relres {rel:.2e} (converged={ok}) in {dt*1e3:9.2f} ms -> {it/dt:8.1f} iterations/s, "
                  f"max|x - x_true| = {err.item():.2e}, powers strategy {ctx.query('last_mpk_strategy')}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
