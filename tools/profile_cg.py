"""Fixed command for an ncu launch list of the two CG variants on the 256^3 operator (device-resident):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/profile_cg.py [--its 16]
Run without ncu it prints the per-iteration wall time of both (the number DESIGN.md section 5 quotes)."""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--its", type=int, default=16)
args = ap.parse_args()
A = matgen.laplace3d_7pt(args.grid)
ctx = nsk.Context(0)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
xt = ctx.to_device(matgen.vec_uniform(A.n, 1))
b = ctx.empty(A.n)
dA.spmv(xt, b)
xs = ctx.empty(A.n)
for s in (1, 4):
    dA.cg(b, xs, tol=1e-300, maxit=args.its, sstep=s)  # warm: plans, workspaces
    ctx.sync()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    _, it, _, _ = dA.cg(b, xs, tol=1e-300, maxit=args.its, sstep=s)
    ctx.sync()
    dt = time.perf_counter() - t0
    print(f"sstep={s}: {it} iterations, {dt / it * 1e3:.4f} ms per iteration, {ctx.launch_count - l0} launches", flush=True)
