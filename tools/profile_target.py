"""Small fixed command for ncu: builds the 256^3 operator, runs a few launches of one kernel family.

    python tools/profile_target.py --what spmv|mpk_wave|mpk_levels [--grid 256] [--reps 3] [--k 4]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="spmv")
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--slack", type=int, default=-1)
ap.add_argument("--wave-variant", type=int, default=0, help="0 = default, n = table entry n-1")
ap.add_argument("--opt", action="append", default=[], help="name=value context option (repeatable)")
ap.add_argument("--mpk-kernel", type=int, default=-1)
args = ap.parse_args()
A = matgen.laplace3d_7pt(args.grid)
ctx = nsk.Context(0)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
x = ctx.to_device(matgen.vec_uniform(A.n, 1))
lv = [ctx.empty(A.n) for _ in range(args.k)]
ctx.set_option("wave_slack_pct", args.slack)
pass  # (wave_l2_pct is left at its default; override with --opt wave_l2_pct=N)
for o in args.opt:
    name, val = o.split("=")
    ctx.set_option(name, int(val))
e0, e1 = ctx.event(), ctx.event()
for i in range(args.reps + 1):
    if i == 1:
        e0.record()
    if args.what == "spmv":
        dA.spmv(x, lv[0])
    elif args.what == "mpk":
        ctx.set_option("mpk_kernel", args.mpk_kernel if args.mpk_kernel >= 0 else 0)
        dA.mpk(args.k, x, lv)
    else:
        ctx.set_option("mpk_kernel", 1)
        dA.mpk(args.k, x, lv)
e1.record()
ctx.sync()
print(f"{args.what} grid={args.grid} k={args.k}: {e0.elapsed_ms(e1)/args.reps:.4f} ms per call, launches={ctx.launch_count}")
