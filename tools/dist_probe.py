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
    t = torch.zeros(world, device=f"cuda:{local}", dtype=torch.float64); t[rank] = ms; dist.all_reduce(t)
    per_rank.append([float(v) for v in t.tolist()])
    return float(t.max().item())
per_rank = []
def plan_info():
    return {q: ctx.query(q) for q in ("sell_uniform_width", "sell_reach", "sell_lead", "sell_grid", "sell_identity_tiles",
                                      "sell_ntiles", "sell_staged", "sell_ngroups")}
res = {}
sx = op.new_vector(shared=True); op.set_owned(sx, np.sin(0.001 * np.arange(op.n_owned)))
res["halo_exchange depth 4, NCCL (plain vector)"] = timed(lambda: op.halo_exchange(dx, 4))
res["halo_exchange depth 1, NCCL (plain vector)"] = timed(lambda: op.halo_exchange(dx, 1))
res["halo_exchange depth 4, push (registered vector)"] = timed(lambda: op.halo_exchange(sx, 4))
res["halo_exchange depth 1, push (registered vector)"] = timed(lambda: op.halo_exchange(sx, 1))
res["mpk k=4, NCCL exchange + fused kernel"] = timed(lambda: op.mpk(4, dx, lv))
res["mpk k=4, push + fused kernel + ack"] = timed(lambda: op.mpk(4, sx, lv))
ctx.set_option("halo_push", -1)
res["mpk k=4, NO exchange (fused kernel on the slab + ghost rows alone)"] = timed(lambda: op.mpk(4, sx, lv))
info_dist = plan_info()
for l2 in (70, 80, 95):
    ctx.set_option("wave_l2_pct", l2)
    res[f"mpk k=4, NO exchange, wave_l2_pct={l2}"] = timed(lambda: op.mpk(4, sx, lv))
ctx.set_option("wave_l2_pct", 0)
ctx.set_option("halo_push", 1)
res["spmv, NCCL exchange depth 1 + product"] = timed(lambda: op.spmv(dx, lv[0]))
res["spmv, push depth 1 + product + ack"] = timed(lambda: op.spmv(sx, lv[0]))
# the same 256^3 operator as a plain single-GPU operator in the same process (no ghost rows, no breaks)
from navierstokes_b200 import matgen
A = matgen.laplace3d_7pt(g)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
px = ctx.to_device(np.sin(0.001 * np.arange(A.n))); plv = [ctx.empty(A.n) for _ in range(4)]
res["mpk k=4, single-GPU operator of the same size (control)"] = timed(lambda: dA.mpk(4, px, plv))
info_single = plan_info()
infos = [None] * world
dist.all_gather_object(infos, info_dist)
if rank == 0:
    print(f"# world={world}, {g}x{g}x{g} per GPU, rows local {op.n_rows_local} owned {op.n_owned} cols {op.n_cols_local}")
    for (k, v), pr in zip(res.items(), per_rank): print(f"{k:70s} {v*1e3:9.1f} us   per rank: " + " ".join(f"{u*1e3:.1f}" for u in pr))
    for r, i in enumerate(infos): print(f"plan, distributed operator, rank {r}:", i)
    print("plan, single-GPU operator:", info_single)
dist.destroy_process_group()
