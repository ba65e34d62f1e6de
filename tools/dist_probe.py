"""Where does a distributed powers step spend its time?  torchrun, one rank per GPU (weak: 256x256x256 per GPU)."""
import datetime, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
import navierstokes_b200 as nsk
from navierstokes_b200 import distributed as nd
ctx = nsk.Context(local)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
op = nd.DistStencil3D(ctx, dist, g, g, g * world, halo_depth=4)
dx = op.new_vector(); op.set_owned(dx, np.sin(0.001 * np.arange(op.n_owned)))
lv = [op.new_vector() for _ in range(4)]
def timed(fn, reps=50):
    for _ in range(5): fn()
    ctx.sync(); dist.barrier(); ctx.sync()
    e0, e1 = ctx.event(), ctx.event(); e0.record()
    for _ in range(reps): fn()
    e1.record(); ms = e0.elapsed_ms(e1) / reps
    t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
res = {}
res["halo_exchange depth 4"] = timed(lambda: op.halo_exchange(dx, 4))
res["halo_exchange depth 1"] = timed(lambda: op.halo_exchange(dx, 1))
res["mpk k=4 (exchange + fused kernel)"] = timed(lambda: op.mpk(4, dx, lv))
res["spmv (exchange depth 1 + product)"] = timed(lambda: op.spmv(dx, lv[0]))
if rank == 0:
    print(f"# world={world}, {g}x{g}x{g} per GPU, rows local {op.n_rows_local} owned {op.n_owned} cols {op.n_cols_local}")
    for k, v in res.items(): print(f"{k:40s} {v*1e3:9.1f} us")
dist.destroy_process_group()
