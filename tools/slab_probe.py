"""One GPU, one process: the operator rank R of a W-rank slab partition would own (ghost rows and all), WITHOUT any
exchange (option halo_push = -1, measurement only) -- to look at the fused kernel on a slab with ghost rows below /
above in isolation (timings, option sweeps, ncu).

    python tools/slab_probe.py --world 2 --rank 1 [--grid 256] [--strong] [--sweep] [--once]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--rank", type=int, default=1)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--strong", action="store_true", help="split grid^3 over the ranks instead of grid^3 per rank")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--once", action="store_true", help="3 warm-up calls + ONE call (for ncu)")
    ap.add_argument("--timing", action="store_true", help="one call with the dependency-warp wait instrumentation (pk_timing)")
    ap.add_argument("--set", action="append", default=[], help="option=value, applied before the first call")
    args = ap.parse_args()
    import navierstokes_b200 as nsk
    from navierstokes_b200 import distributed as nd
    g, W, R, K = args.grid, args.world, args.rank, args.k
    nz = g if args.strong else g * W
    starts = nd.slab_row_starts(nz, g * g, W)
    prov = nd.StencilProvider(g, g, nz)
    plans = [nd.Plan.build(W, r, starts, K, prov) for r in range(W)]
    reqs = []
    for p in plans:
        mine = {q: p.requests(q) for q in range(W) if q != p.rank}
        reqs.append({q: v for q, v in mine.items() if len(v[0])})
    plan = plans[R]
    plan.exchange_requests(all_requests=reqs)
    ctx = nsk.Context(0)
    for s in args.set:
        name, v = s.split("=")
        ctx.set_option(name, int(v))
    h = C.c_void_p()
    ctx._ck(ctx.lib.nsk_csr_create_dist(ctx.h, plan.h, C.byref(h)))
    ctx.set_option("halo_push", -1)
    x = ctx.zeros(plan.n_cols_local)
    xh = np.sin(0.001 * np.arange(plan.n_cols_local))
    ctx._ck(ctx.lib.nsk_memcpy(ctx.h, x.ptr, C.c_void_p(xh.ctypes.data), 8 * plan.n_cols_local, 0))
    ctx.sync()
    lv = [ctx.zeros(plan.n_cols_local) for _ in range(K)]
    ptrs = (C.c_void_p * K)(*[l.ptr.value for l in lv])

    def call():
        ctx._ck(ctx.lib.nsk_mpk(h, K, x.ptr, ptrs, 0, 1))

    def timed(reps=30):
        for _ in range(4):
            call()
        ctx.sync()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        return e0.elapsed_ms(e1) / reps * 1e3

    print(f"# rank {R} of {W}, {g}x{g}x{nz} grid, owned {plan.n_owned} rows, local rows {plan.n_rows_local}, cols {plan.n_cols_local}, "
          f"ring_start {list(plan.ring_start)}", flush=True)
    if args.once:
        for _ in range(3):
            call()
        ctx.sync()
        call()
        ctx.sync()
        print("one call done; launches", ctx.launch_count)
        return
    print(f"k={K} default: {timed():.1f} us  strategy {ctx.query('last_mpk_strategy')}  plan "
          + str({q: ctx.query(q) for q in ('sell_reach', 'sell_lead', 'sell_grid', 'sell_ntiles', 'sell_ngroups')}), flush=True)
    if args.timing:
        ctx.set_option("pk_timing", 1)
        call()
        ctx.sync()
        ctx.set_option("pk_timing", 0)
    if args.sweep:
        for name, vals in (("wave_l2_pct", (80, 88, 92, 96, 100, 105)), ("sell_chunk", (2, 3, 4)), ("pipe_w0_pct", (90, 110, 125)),
                           ("sell_pf_dist", (1, 3, 4)), ("pipe_interleave", (0,))):
            for v in vals:
                ctx.set_option(name, v)
                print(f"  {name}={v}: {timed():.1f} us", flush=True)
            ctx.set_option(name, 1 if name == "pipe_interleave" else 0)


if __name__ == "__main__":
    main()
