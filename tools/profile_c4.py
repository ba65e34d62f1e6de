"""Unstructured operator (BASELINE config 4: tet P1 + RCM) through the fused explicit-column kernel: variants side by side.

    python tools/profile_c4.py [--m 200] [--k 4] [--variants 0,16,17] [--timing] [--ncu]

--ncu: one warm call + one call per variant only (the command to put under `ncu -k regex:sell_tma --set full`).
Every variant is compared bit for bit with k products.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402
from bench_c4 import tetgen, timed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=200)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--variants", default="0,8,17", help="sell_tma values: 0 = default (16 entries per trip, 2 CTAs/SM), 8, 17")
ap.add_argument("--ctas", default="0", help="comma list of sell_ctas_per_sm values (0 = occupancy)")
ap.add_argument("--chunks", default="0", help="comma list of sell_chunk values (0 = default)")
ap.add_argument("--l2", default="0", help="comma list of wave_l2_pct values (0 = default 88)")
ap.add_argument("--flags", default="-1", help="comma list of sell_flags values (-1 = default 3)")
ap.add_argument("--timing", action="store_true")
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()

t0 = time.time()
A = matgen.rcm_reorder(tetgen(args.m, 2, 32))
print(f"# tet P1 {args.m + 1}^3 nodes: n={A.nrows} nnz={A.nnz} bandwidth={matgen.bandwidth(A)} generated + RCM in {time.time() - t0:.1f}s", flush=True)
ctx = nsk.Context(0)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
n, k = A.nrows, args.k
dx = ctx.to_device(matgen.vec_uniform(n, seed=1))
lv = [ctx.empty(n) for _ in range(k)]
ctx.set_option("mpk_kernel", 1)
dA.mpk(k, dx, lv)
ref = [l.to_host() for l in lv]
if not args.ncu:
    ms1 = timed(ctx, lambda: dA.mpk(k, dx, lv), args.reps)
    print(f"k={k} as {k} products: {ms1:.4f} ms  {n * k / ms1 / 1e3:.1f} rows/us", flush=True)
ctx.set_option("mpk_kernel", 5)
for v in [int(x) for x in args.variants.split(",")]:
    for c in [int(x) for x in args.ctas.split(",")]:
      for ch in [int(x) for x in args.chunks.split(",")]:
       for l2 in [int(x) for x in args.l2.split(",")]:
        for fl in [int(x) for x in args.flags.split(",")]:
            ctx.set_option("sell_tma", v)
            ctx.set_option("sell_ctas_per_sm", c)
            ctx.set_option("sell_chunk", ch)
            ctx.set_option("wave_l2_pct", l2)
            ctx.set_option("sell_flags", fl)
            for l in lv:
                ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * n)
            l0 = ctx.launch_count
            dA.mpk(k, dx, lv)
            ctx.sync()
            launches = ctx.launch_count - l0
            same = all(np.array_equal(lv[i].to_host().view(np.int64), ref[i].view(np.int64)) for i in range(k))
            if args.ncu:
                print(f"variant sell_tma={v} ctas={c} chunk={ch}: launches={launches} bitwise={same}", flush=True)
                continue
            ms = timed(ctx, lambda: dA.mpk(k, dx, lv), args.reps)
            print(json.dumps({"sell_tma": v, "ctas_per_sm": c, "chunk": ch, "l2_pct": l2, "flags": fl, "k": k, "ms": round(ms, 4), "rows_per_us": round(n * k / ms / 1e3, 1),
                              "vs_products": round(ms1 / ms, 3), "launches": launches, "strategy": ctx.query("last_mpk_strategy"),
                              "grid": ctx.query("sell_grid"), "reach": ctx.query("sell_reach"), "lead": ctx.query("sell_lead"),
                              "bitwise": bool(same)}), flush=True)
            if args.timing:
                ctx.set_option("pk_timing", 1)
                dA.mpk(k, dx, lv)
                ctx.sync()
                ctx.set_option("pk_timing", 0)
