"""GPU sweep of the SpMV kernel variants on the BASELINE configs (device-resident, CUDA-event timed).

usage: python tools/sweep_spmv.py [--cfg c3|c2|fem] [--reps 20]
Prints one line per (kernel, variant, ctas/SM): ms, GB/s (algorithmic bytes), fraction of the measured
HBM copy peak.  Every timed configuration is first checked bit-for-bit against the simple kernel.
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="c3")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--modes", default="0")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--wave-variants", default="0,1,2,3,4,5,6,7,8,9,10,11")
    ap.add_argument("--slacks", default="50,100,150")
    ap.add_argument("--skip-spmv", action="store_true")
    ap.add_argument("--wave-static", type=int, default=0)
    ap.add_argument("--pipe-variants", default="0,1,2,3,4,5,6,7,8,9")
    ap.add_argument("--pipe-leads", default="50,100,200")
    ap.add_argument("--pipe-interleave", default="1,0")
    ap.add_argument("--ks", default="4")
    ap.add_argument("--packed-variants", default="0,1,2,3,4,5,6,7,8,9")
    ap.add_argument("--bp-global", type=int, default=0)
    ap.add_argument("--w0", default="100")
    args = ap.parse_args()
    t0 = time.time()
    if args.cfg == "c3":
        A = matgen.laplace3d_7pt(256)
    elif args.cfg == "c2":
        A = matgen.laplace2d_5pt(4096)
    elif args.cfg == "fem":
        A = matgen.fem_baij4(40)
    elif "x" in args.cfg:
        A = matgen.laplace3d_7pt(*[int(v) for v in args.cfg.split("x")])
    else:
        A = matgen.laplace3d_7pt(int(args.cfg))
    print(f"# {args.cfg}: n={A.n} nnz={A.nnz} built in {time.time()-t0:.1f}s", flush=True)
    ctx = nsk.Context(0)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    y = ctx.empty(A.n)
    yref = ctx.empty(A.n)
    peak = peak_gbs()
    B = dA.spmv_bytes
    for mode in ([] if args.skip_spmv else [int(m) for m in args.modes.split(",")]):
        ctx.set_option("spmv_kernel", 1)
        dA.spmv(x, yref, mode)
        ref = yref.to_host()
        ms = timed(ctx, lambda: dA.spmv(x, y, mode), args.reps)
        print(f"mode={mode} simple            : {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s {B/ms/1e6/peak:6.3f} of measured peak", flush=True)
        ctx.set_option("spmv_kernel", 2)
        for var in ((3, 4) if args.quick else range(1, 9)):
            ctx.set_option("stream_variant", var)
            for cps in ((0,) if args.quick else (0, 1, 2, 3, 4)):
                ctx.set_option("spmv_ctas_per_sm", cps)
                try:
                    dA.spmv(x, y, mode)
                    got = y.to_host()
                except nsk.NskError as e:
                    print(f"mode={mode} stream v{var-1} cps={cps}: ERROR {e}", flush=True)
                    raise
                same = np.array_equal(got.view(np.int64), ref.view(np.int64)) if mode != 2 else bool(
                    np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref))
                ms = timed(ctx, lambda: dA.spmv(x, y, mode), args.reps)
                print(f"mode={mode} stream v{var-1} cps={cps or 'dflt'} : {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s "
                      f"{B/ms/1e6/peak:6.3f} of measured peak  {'OK' if same else 'MISMATCH'}", flush=True)
        ctx.set_option("spmv_ctas_per_sm", 0)
        ctx.set_option("stream_variant", 0)
    # matrix powers
    for k in [int(v) for v in args.ks.split(',')]:
        lv = [ctx.empty(A.n) for _ in range(k)]
        ctx.set_option("mpk_kernel", 1)
        dA.mpk(k, x, lv, 0)
        ref = [l.to_host() for l in lv]
        ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 4))
        Bk = dA.mpk_bytes(k)
        print(f"mpk k={k} levels           : {ms:8.4f} ms  B_mpk rate {Bk/ms/1e6:8.1f} GB/s ({Bk/ms/1e6/peak:5.3f}), "
              f"SpMV-equivalent {k*B/ms/1e6:8.1f} GB/s", flush=True)
        ctx.set_option("mpk_kernel", 2)
        ctx.set_option("wave_static", args.wave_static)
        ctx.set_option("wave_l2_pct", 400)  # never refuse in the sweep: we want to see the cliff
        for wv in [int(v) for v in args.wave_variants.split(",") if v != ""]:
            ctx.set_option("wave_variant", wv + 1)
            for slack in [int(v) for v in args.slacks.split(",")]:
                ctx.set_option("wave_slack_pct", slack)
                for l in lv:
                    ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * A.n)
                l0 = ctx.launch_count
                dA.mpk(k, x, lv, 0)
                fused = (ctx.launch_count - l0) == 1
                same = all(np.array_equal(lv[i].to_host().view(np.int64), ref[i].view(np.int64)) for i in range(k))
                ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 4))
                print(f"mpk k={k} wavefront v{wv} slack={slack:3d}% {'fused' if fused else 'LEVELS'}: {ms:8.4f} ms  B_mpk rate "
                      f"{Bk/ms/1e6:8.1f} GB/s ({Bk/ms/1e6/peak:5.3f}), SpMV-equivalent {k*B/ms/1e6:8.1f} GB/s  "
                      f"{'OK' if same else 'MISMATCH'}", flush=True)
        ctx.set_option("mpk_kernel", 3)
        for pv in [int(v) for v in args.pipe_variants.split(",") if v != ""]:
            ctx.set_option("pipe_variant", pv + 1)
            for il in [int(v) for v in args.pipe_interleave.split(",")]:
                ctx.set_option("pipe_interleave", il)
                for lead in [int(v) for v in args.pipe_leads.split(",")]:
                    ctx.set_option("wave_slack_pct", lead)
                    for l in lv:
                        ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * A.n)
                    l0 = ctx.launch_count
                    dA.mpk(k, x, lv, 0)
                    fused = (ctx.launch_count - l0) == 1
                    same = all(np.array_equal(lv[i].to_host().view(np.int64), ref[i].view(np.int64)) for i in range(k))
                    ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 4))
                    print(f"mpk k={k} pipeline v{pv} il={il} lead={lead:3d}% {'fused' if fused else 'LEVELS'}: {ms:8.4f} ms  B_mpk rate "
                          f"{Bk/ms/1e6:8.1f} GB/s ({Bk/ms/1e6/peak:5.3f}), SpMV-equivalent {k*B/ms/1e6:8.1f} GB/s  "
                          f"{'OK' if same else 'MISMATCH'}", flush=True)
        ctx.set_option("pipe_variant", 0)
        ctx.set_option("mpk_kernel", 4)
        ctx.set_option("pipe_bp_global", args.bp_global)
        for pv in [int(v) for v in args.packed_variants.split(",") if v != ""]:
            ctx.set_option("packed_variant", pv + 1)
            t1 = time.time()
            pb = dA.packed_bytes
            print(f"# packed v{pv}: {pb/1e6:.1f} MB ({pb/max(1,A.nnz):.2f} B/nnz) packed in {time.time()-t1:.2f}s", flush=True)
            ctx.set_option("spmv_kernel", 3)
            dA.spmv(x, y, 0)
            same = np.array_equal(y.to_host().view(np.int64), ref[0].view(np.int64))
            ms = timed(ctx, lambda: dA.spmv(x, y, 0), args.reps)
            print(f"spmv packed v{pv}: {ms:8.4f} ms {B/ms/1e6:8.1f} GB/s {B/ms/1e6/peak:6.3f} of measured peak  {'OK' if same else 'MISMATCH'}", flush=True)
            ctx.set_option("spmv_kernel", 0)
            for il, w0 in [(int(v), int(w)) for v in args.pipe_interleave.split(",") for w in args.w0.split(",")]:
                ctx.set_option("pipe_interleave", il)
                ctx.set_option("pipe_w0_pct", w0)
                for lead in [int(v) for v in args.pipe_leads.split(",")]:
                    ctx.set_option("wave_slack_pct", lead)
                    ctx.set_option("wave_l2_pct", 400 if lead >= 0 else 0)
                    for l in lv:
                        ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * A.n)
                    l0 = ctx.launch_count
                    dA.mpk(k, x, lv, 0)
                    fused = (ctx.launch_count - l0) == 1
                    same = all(np.array_equal(lv[i].to_host().view(np.int64), ref[i].view(np.int64)) for i in range(k))
                    ms = timed(ctx, lambda: dA.mpk(k, x, lv, 0), max(4, args.reps // 4))
                    print(f"mpk k={k} packed v{pv} il={il} w0={w0} lead={lead:3d}% {'fused' if fused else 'LEVELS'}: {ms:8.4f} ms  B_mpk rate "
                          f"{Bk/ms/1e6:8.1f} GB/s ({Bk/ms/1e6/peak:5.3f}), SpMV-equivalent {k*B/ms/1e6:8.1f} GB/s  "
                          f"{'OK' if same else 'MISMATCH'}", flush=True)
        ctx.set_option("packed_variant", 0)
        ctx.set_option("pipe_w0_pct", 0)
        ctx.set_option("pipe_interleave", 1)
        ctx.set_option("wave_slack_pct", -1)
        ctx.set_option("wave_l2_pct", 0)
        ctx.set_option("spmv_ctas_per_sm", 0)
        ctx.set_option("wave_variant", 0)


if __name__ == "__main__":
    main()
