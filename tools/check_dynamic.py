"""Dynamic-claim instances of the packed kernel (option pk_flags, bit 8) against the static ones: first bit-for-bit on
small and odd-shaped operators (one and two right-hand sides, k = 2..5, both exact modes), then timing on 256^3.

    timeout 600 python tools/check_dynamic.py [--quick] [--timing]

The dynamic instances were written without a GPU at hand (round 1 ran out of GPU minutes): run this BEFORE making them
the default.  Exit code 0 = every comparison was bit-identical.  Each launch traps instead of hanging when the
protocol is broken (bounded spins), so a failure shows up as a CUDA error, not as a stuck box.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402

STATIC, DYNAMIC = 1, 1 | 8
# claim-ahead depth of the dynamic kernels: default 1 item beyond the stage ring, flag 16 -> 2, flag 32 -> 0
DYN_MODES = (("ahead 1", 1 | 8), ("ahead 2", 1 | 8 | 16), ("ahead 0", 1 | 8 | 32))


def run(ctx, dA, k, xs, flags, mode):
    ctx.set_option("pk_flags", flags)
    if len(xs) == 1:
        lv = dA.mpk(k, xs[0], mode=mode)
        out = [[v.to_host() for v in lv]]
    else:
        lv = dA.mpk_multi(k, xs, mode=mode)
        out = [[v.to_host() for v in row] for row in lv]
    strategy = ctx.query("last_mpk_strategy")
    return out, strategy


def parity(ctx, quick):
    ops = [("7pt 64x24x20", matgen.laplace3d_7pt(64, 24, 20)),
           ("7pt 61x17x23 (odd n)", matgen.laplace3d_7pt(61, 17, 23)),
           ("5pt 300x41", matgen.laplace2d_5pt(300, 41)),
           ("7pt 128x64x48", matgen.laplace3d_7pt(128, 64, 48))]
    if not quick:
        ops += [("random stencil %d" % s, matgen.random_stencil3d(40, 24, 16, seed=s, max_points=6)) for s in range(3)]
        ops += [("7pt 256x128x64", matgen.laplace3d_7pt(256, 128, 64))]
    bad = 0
    ctx.set_option("wave_l2_pct", 1000)  # small operators: let the whole window count as L2-resident so that they fuse
    for name, A in ops:
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        for nv in (1, 2):
            xs = [ctx.to_device(matgen.vec_uniform(A.n, 3 + v)) for v in range(nv)]
            for k in (2, 3, 4, 5):
                for mode in (nsk.EXACT_FMA, nsk.EXACT_MULADD):
                    ref, s0 = run(ctx, dA, k, xs, STATIC, mode)
                    for rep in range(2 if quick else 4):  # the claim order differs from run to run: repeat
                        got, s1 = run(ctx, dA, k, xs, DYNAMIC, mode)
                        same = all(np.array_equal(a.view(np.uint64), b.view(np.uint64))
                                   for ra, rb in zip(ref, got) for a, b in zip(ra, rb))
                        if not same or s0 != s1:
                            bad += 1
                            print(f"MISMATCH {name} nv={nv} k={k} mode={mode} rep={rep} strategies {s0}/{s1}", flush=True)
        print(f"{name:28s} n={A.n:8d} packed={'yes' if dA.packed_bytes else 'NO (CSR kernels: nothing compared)'} "
              f"strategy {s0}  {'ok' if bad == 0 else 'FAILED so far: %d' % bad}", flush=True)
        dA.close()
    ctx.set_option("wave_l2_pct", 0)
    return bad


def timed(ctx, fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / reps


def timing(ctx, show_cycle):
    A = matgen.laplace3d_7pt(256)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    x2 = ctx.to_device(matgen.vec_uniform(A.n, 2))
    for k in (2, 3, 4):
        lv = [ctx.empty(A.n) for _ in range(k)]
        lv2 = [[ctx.empty(A.n) for _ in range(k)] for _ in range(2)]
        for l2 in (0, 60, 80):  # 0 = default budget (70 %)
            ctx.set_option("wave_l2_pct", l2)
            row = []
            for flags in (STATIC, DYNAMIC):
                ctx.set_option("pk_flags", flags)
                t1 = timed(ctx, lambda: dA.mpk(k, x, lv))
                t2 = timed(ctx, lambda: dA.mpk_multi(k, [x, x2], lv2))
                row.append((t1, t2))
            print(f"256^3 k={k} L2 budget {l2 or 70:3d} %: one vector static {row[0][0]:.4f} ms, dynamic {row[1][0]:.4f} ms | "
                  f"two vectors static {row[0][1]:.4f} ms, dynamic {row[1][1]:.4f} ms", flush=True)
        ctx.set_option("wave_l2_pct", 0)
    if show_cycle:
        lv = [ctx.empty(A.n) for _ in range(4)]
        for flags in (STATIC, DYNAMIC):
            ctx.set_option("pk_flags", flags)
            dA.mpk(4, x, lv)
            ctx.set_option("pk_timing", 1)
            print(f"# stage cycle, pk_flags={flags}", file=sys.stderr, flush=True)
            dA.mpk(4, x, lv)
            ctx.sync()
            ctx.set_option("pk_timing", 0)
    ctx.set_option("pk_flags", STATIC)


def fast(ctx):
    """The whole check in well under two minutes: a few small operators, then 256^3 compared bit for bit and timed."""
    bad = 0
    ctx.set_option("wave_l2_pct", 1000)
    for name, A in (("7pt 64x24x20", matgen.laplace3d_7pt(64, 24, 20)), ("7pt 61x17x23 (odd n)", matgen.laplace3d_7pt(61, 17, 23)),
                    ("7pt 128x64x48", matgen.laplace3d_7pt(128, 64, 48))):
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        for nv in (1, 2):
            xs = [ctx.to_device(matgen.vec_uniform(A.n, 3 + v)) for v in range(nv)]
            for k in (2, 4):
                ref, s0 = run(ctx, dA, k, xs, STATIC, nsk.EXACT_FMA)
                for rep in range(2):
                    for _, flags in DYN_MODES:
                        got, s1 = run(ctx, dA, k, xs, flags, nsk.EXACT_FMA)
                        same = all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for ra, rb in zip(ref, got) for a, b in zip(ra, rb))
                        bad += (not same) or s0 != s1
                print(f"{name} nv={nv} k={k}: strategy {s0}/{s1} {'ok' if bad == 0 else 'MISMATCH'}", flush=True)
        dA.close()
    ctx.set_option("wave_l2_pct", 0)
    if bad:
        return bad
    A = matgen.laplace3d_7pt(256)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    x2 = ctx.to_device(matgen.vec_uniform(A.n, 2))
    print("256^3 ready", flush=True)
    for k in (4, 2, 3):
        lv = [ctx.empty(A.n) for _ in range(k)]
        ctx.set_option("pk_flags", STATIC)
        dA.mpk(k, x, lv)
        ref = [v.to_host() for v in lv]
        same = True
        for _, flags in DYN_MODES:
            ctx.set_option("pk_flags", flags)
            for v in lv:
                ctx.axpy(1.0, x, v)  # spoil the previous result: a launch that wrote nothing must not pass
            dA.mpk(k, x, lv)
            same = same and all(np.array_equal(a.view(np.uint64), v.to_host().view(np.uint64)) for a, v in zip(ref, lv))
        bad += not same
        ts = []
        for flags in (STATIC,) + tuple(f for _, f in DYN_MODES) + (STATIC,):
            ctx.set_option("pk_flags", flags)
            ts.append(timed(ctx, lambda: dA.mpk(k, x, lv), reps=20))
        print(f"256^3 k={k}: {'bit-identical' if same else 'MISMATCH'}; static {ts[0]:.4f} / {ts[4]:.4f} ms, dynamic ahead 1 {ts[1]:.4f}, "
              f"ahead 2 {ts[2]:.4f}, ahead 0 {ts[3]:.4f} ms (strategy {ctx.query('last_mpk_strategy')})", flush=True)
    k = 4
    lv2 = [[ctx.empty(A.n) for _ in range(k)] for _ in range(2)]
    ts = []
    for flags in (STATIC,) + tuple(f for _, f in DYN_MODES):
        ctx.set_option("pk_flags", flags)
        ts.append(timed(ctx, lambda: dA.mpk_multi(k, [x, x2], lv2), reps=10))
    print(f"256^3 k=4, two vectors: static {ts[0]:.4f} ms, dynamic ahead 1 {ts[1]:.4f}, ahead 2 {ts[2]:.4f}, ahead 0 {ts[3]:.4f} ms", flush=True)
    lv4 = [ctx.empty(A.n) for _ in range(4)]
    for l2 in (60, 80, 90):
        ctx.set_option("wave_l2_pct", l2)
        ts = []
        for flags in (STATIC,) + tuple(f for _, f in DYN_MODES):
            ctx.set_option("pk_flags", flags)
            ts.append(timed(ctx, lambda: dA.mpk(4, x, lv4), reps=10))
        print(f"256^3 k=4, L2 budget {l2} %: static {ts[0]:.4f} ms, dynamic ahead 1 {ts[1]:.4f}, ahead 2 {ts[2]:.4f}, ahead 0 {ts[3]:.4f} ms", flush=True)
    ctx.set_option("wave_l2_pct", 0)
    lv = [ctx.empty(A.n) for _ in range(4)]
    for flags in (STATIC,) + tuple(f for _, f in DYN_MODES):
        ctx.set_option("pk_flags", flags)
        dA.mpk(4, x, lv)
        ctx.set_option("pk_timing", 1)
        print(f"# stage cycle, pk_flags={flags}", file=sys.stderr, flush=True)
        dA.mpk(4, x, lv)
        ctx.sync()
        ctx.set_option("pk_timing", 0)
    ctx.set_option("pk_flags", STATIC)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--fast", action="store_true", help="short parity + 256^3 parity and timing, under two minutes")
    ap.add_argument("--timing", action="store_true", help="also time 256^3 (k = 2, 3, 4) and print the stage cycle of both")
    args = ap.parse_args()
    ctx = nsk.Context(0)
    if args.fast:
        bad = fast(ctx)
        print("fast check:", "all bit-identical" if bad == 0 else f"{bad} mismatches", flush=True)
        return 1 if bad else 0
    bad = parity(ctx, args.quick)
    print("parity:", "all bit-identical" if bad == 0 else f"{bad} mismatches", flush=True)
    if bad == 0 and args.timing:
        timing(ctx, True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
