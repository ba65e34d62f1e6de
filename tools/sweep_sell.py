"""GPU sweep of the sliced-ELL kernel (csrc/sell.cu) against the packed kernel on a BASELINE config (device-resident,
CUDA-event timed, L2-sized rotation not needed: the operator is far larger than L2).

usage: python tools/sweep_sell.py [--cfg c3|c2|tet:M|fem:M|NXxNYxNZ] [--k 4] [--reps 12] [--chunks 1,2,4] [--cps 0,2] ...
Every timed configuration is first checked bit-for-bit against k launches of the CSR streaming kernel.
"""
import argparse
import itertools
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402


def peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    return json.loads(p.read_text())["hbm_gbs"] if p.exists() else 6650.0


def timed(ctx, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / reps


def ints(s):
    return [int(v) for v in s.split(",") if v != ""]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="c3")
    ap.add_argument("--k", default="4")
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--chunks", default="1,2,3")
    ap.add_argument("--cps", default="0")
    ap.add_argument("--flags", default="3")
    ap.add_argument("--pf", default="2")
    ap.add_argument("--l2", default="0")
    ap.add_argument("--w0", default="0")
    ap.add_argument("--spmv-chunks", default="1,2,3")
    ap.add_argument("--stream", default="0")
    ap.add_argument("--rows", default="0")
    ap.add_argument("--tma", default="0")
    ap.add_argument("--geom", default="0", help="sell_geom: 0 default, 2 = masked consumer path everywhere (A/B of the straight-line path)")
    ap.add_argument("--no-packed", action="store_true")
    args = ap.parse_args()
    t0 = time.time()
    if args.cfg == "c3":
        A = matgen.laplace3d_7pt(256)
    elif args.cfg == "c2":
        A = matgen.laplace2d_5pt(4096)
    elif args.cfg.startswith("tet:"):
        A = matgen.tet_p1_laplacian(int(args.cfg[4:]), permute_seed=2, rcm=True)
    elif args.cfg.startswith("fem:"):
        A = matgen.fem_baij4(int(args.cfg[4:]))
    else:
        A = matgen.laplace3d_7pt(*[int(v) for v in args.cfg.split("x")])
    print(f"# {args.cfg}: n={A.n} nnz={A.nnz} built in {time.time()-t0:.1f}s", flush=True)
    ctx = nsk.Context(0)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    y = ctx.empty(A.n)
    peak = peak_gbs()
    B = dA.spmv_bytes
    same_bits = lambda a, b: np.array_equal(a.view(np.int64), b.view(np.int64))

    # reference bits: the CSR streaming kernel
    ctx.set_option("spmv_kernel", 2)
    dA.spmv(x, y, 0)
    ref1 = y.to_host()
    ms = timed(ctx, lambda: dA.spmv(x, y, 0), args.reps)
    print(f"spmv stream (CSR)       : {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s {B/ms/1e6/peak:6.3f} of measured peak", flush=True)
    if not args.no_packed:
        ctx.set_option("spmv_kernel", 3)
        dA.spmv(x, y, 0)
        ok = same_bits(y.to_host(), ref1)
        ms = timed(ctx, lambda: dA.spmv(x, y, 0), args.reps)
        print(f"spmv packed (kernel {ctx.query('last_spmv_kernel')})  : {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s {B/ms/1e6/peak:6.3f}  "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
    ctx.set_option("spmv_kernel", 4)
    t1 = time.time()
    dA.spmv(x, y, 0)
    ctx.sync()
    print(f"# sliced-ELL setup + first product {time.time()-t1:.2f}s", flush=True)
    for chunk, cps, stream, rows, tma, geom in itertools.product(ints(args.spmv_chunks), ints(args.cps), ints(args.stream),
                                                                 ints(args.rows), ints(args.tma), ints(args.geom)):
        ctx.set_option("sell_geom", geom)
        ctx.set_option("sell_tma", tma)
        ctx.set_option("sell_rows", rows)
        ctx.set_option("sell_chunk", chunk)
        ctx.set_option("sell_stream", stream)
        ctx.set_option("sell_ctas_per_sm", cps)
        ctx.lib.nsk_memset0(ctx.h, y.ptr, 8 * A.n)
        dA.spmv(x, y, 0)
        ok = same_bits(y.to_host(), ref1) and ctx.query("last_spmv_kernel") == 4
        ms = timed(ctx, lambda: dA.spmv(x, y, 0), args.reps)
        print(f"spmv sell chunk={chunk} cps={cps} stream={stream} rows={rows} tma={tma} geom={geom}: {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s {B/ms/1e6/peak:6.3f}  "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
    ctx.set_option("spmv_kernel", 0)

    for k in ints(args.k):
        lv = [ctx.empty(A.n) for _ in range(k)]
        Bk = dA.mpk_bytes(k)
        ctx.set_option("mpk_kernel", 1)
        ctx.set_option("spmv_kernel", 2)
        dA.mpk(k, x, lv, 0)
        ref = [l.to_host() for l in lv]
        ctx.set_option("spmv_kernel", 0)
        ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 2))
        print(f"mpk k={k} as {k} products: {ms:8.4f} ms  B_mpk {Bk/ms/1e6:8.1f} GB/s ({Bk/ms/1e6/peak:5.3f})", flush=True)
        if not args.no_packed:
            ctx.set_option("mpk_kernel", 4)
            for l in lv:
                ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * A.n)
            dA.mpk(k, x, lv, 0)
            ok = all(same_bits(lv[i].to_host(), ref[i]) for i in range(k))
            ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 2))
            print(f"mpk k={k} packed (strategy {ctx.query('last_mpk_strategy')}): {ms:8.4f} ms  B_mpk {Bk/ms/1e6:8.1f} GB/s "
                  f"({Bk/ms/1e6/peak:5.3f})  {'OK' if ok else 'MISMATCH'}", flush=True)
        ctx.set_option("mpk_kernel", 5)
        for chunk, cps, flags, pf, l2, w0, stream, rows, tma, geom in itertools.product(
                ints(args.chunks), ints(args.cps), ints(args.flags), ints(args.pf), ints(args.l2), ints(args.w0),
                ints(args.stream), ints(args.rows), ints(args.tma), ints(args.geom)):
            ctx.set_option("sell_geom", geom)
            ctx.set_option("sell_tma", tma)
            ctx.set_option("sell_rows", rows)
            ctx.set_option("sell_chunk", chunk)
            ctx.set_option("sell_stream", stream)
            ctx.set_option("sell_ctas_per_sm", cps)
            ctx.set_option("sell_flags", flags)
            ctx.set_option("sell_pf_dist", pf)
            ctx.set_option("wave_l2_pct", l2)
            ctx.set_option("pipe_w0_pct", w0)
            for l in lv:
                ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * A.n)
            l0 = ctx.launch_count
            dA.mpk(k, x, lv, 0)
            nl = ctx.launch_count - l0
            ok = all(same_bits(lv[i].to_host(), ref[i]) for i in range(k))
            ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 2))
            print(f"mpk k={k} sell chunk={chunk} stream={stream} rows={rows} tma={tma} geom={geom} cps={cps} flags={flags} pf={pf} l2={l2} w0={w0} "
                  f"launches={nl} strategy={ctx.query('last_mpk_strategy')}: {ms:8.4f} ms  B_mpk {Bk/ms/1e6:8.1f} GB/s "
                  f"({Bk/ms/1e6/peak:5.3f})  {'OK' if ok else 'MISMATCH'}  reach={ctx.query('sell_reach')} lead={ctx.query('sell_lead')} "
                  f"grid={ctx.query('sell_grid')} groups={ctx.query('sell_ngroups')}", flush=True)
        ctx.set_option("mpk_kernel", 0)


if __name__ == "__main__":
    main()
