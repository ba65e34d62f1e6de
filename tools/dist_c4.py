"""BASELINE config 4 across ranks: RCM-ordered tet-P1 Laplacian cut into contiguous row blocks, depth-k ghost rings
(GlobalCsrProvider), ONE halo exchange per matrix-powers call, ghost levels recomputed on shrinking prefixes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/dist_c4.py [--mesh 200] [--k 8]

Every rank generates the same global operator on the host (fast generator + nsk_rcm), builds its slab, runs k products and the
automatic strategy, and compares its owned rows of ALL k levels bit for bit with the CPU oracle (k x SpMV_CSR_FMA restated) on
the global operator.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=200, dest="m")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import navierstokes_b200 as nsk
    from navierstokes_b200 import matgen, distributed as nd
    from bench_c4 import tetgen
    import oracle
    oracle.build(ref=False)
    k = args.k
    t0 = time.time()
    A = matgen.rcm_reorder(tetgen(args.m, 2, max(1, 32 // world)))
    gen_s = time.time() - t0
    n = A.nrows
    ctx = nsk.Context(local)
    starts = (np.arange(world + 1, dtype=np.int64) * n // world).astype(np.int32)
    t0 = time.time()
    op = nd.DistOperator(ctx, dist, starts, nd.GlobalCsrProvider(A), k)
    plan_s = time.time() - t0
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    x = matgen.vec_uniform(n, seed=1)
    dx = op.new_vector(shared=True)  # halo pushed over NVLink peer memory
    op.set_owned(dx, x[lo:hi])
    lv = [op.new_vector() for _ in range(k)]

    def timed(fn):
        for _ in range(2):
            fn()
        ctx.sync(); dist.barrier()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        t = torch.tensor([e0.elapsed_ms(e1) / args.reps], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"config": f"C4 tet P1 Laplacian {args.m + 1}^3 nodes + nsk_rcm over {world} GPUs, contiguous row blocks, depth-{k} halo, k={k}",
           "n": n, "nnz": int(A.nnz), "generate_rcm_s": round(gen_s, 1), "plan_s": round(plan_s, 1),
           "rank0": {"n_owned": op.n_owned, "n_rows_local": op.n_rows_local, "n_cols_local": op.n_cols_local,
                     "peers": int(ctx.lib.nsk_dist_peer_count(op.h))}}
    res = {}
    for name, strat in (("k_products", 1), ("auto", 0)):
        ctx.set_option("mpk_kernel", strat)
        l0 = ctx.launch_count
        op.mpk(k, dx, lv)
        ctx.sync()
        launches = ctx.launch_count - l0
        got = [op.get_owned(v) for v in lv]
        ms = timed(lambda: op.mpk(k, dx, lv))
        res[name] = {"ms": ms, "launches_per_call": launches, "strategy": ctx.query("last_mpk_strategy"), "levels": got}
    ctx.set_option("mpk_kernel", 0)
    hx = timed(lambda: op.halo_exchange(dx, k))
    # parity: owned rows of all k levels against the oracle on the GLOBAL operator
    ref = oracle.lib.mpk(A.ptrow, A.indcol, A.coef, k, x)
    bad = 0
    for name in res:
        for l in range(k):
            bad += int(np.count_nonzero(res[name]["levels"][l].view(np.int64) != ref[l][lo:hi].view(np.int64)))
        del res[name]["levels"]
    t = torch.tensor([float(bad)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t)
    out.update(res)
    out["halo_exchange_us"] = hx * 1e3
    out["speedup_auto_vs_products"] = res["k_products"]["ms"] / res["auto"]["ms"]
    out["parity"] = {"entries_differing_all_ranks": int(t.item()), "bitwise": bool(t.item() == 0),
                     "against": f"oracle k x SpMV_CSR_FMA on the global operator, owned rows of all {k} levels, every rank"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    op.close()
    dist.destroy_process_group()
    return int(t.item() != 0)


if __name__ == "__main__":
    sys.exit(main())
