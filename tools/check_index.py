"""Index compression of the packed operator (option packed_index, blob format 1) against the plain packed format:
bit-for-bit on small and odd-shaped operators (SpMV, fused powers, one and two right-hand sides, all arithmetic modes),
then 256^3 compared and timed.

    timeout 300 python tools/check_index.py

Exit code 0 = every comparison was bit-identical.  Round-1 result (profiles/r01_check_index*.txt): identical bits,
but the consumer loop of the format-1 instances is slower than the explicit one -- the option stays off.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import matgen  # noqa: E402


def same_bits(a, b):
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


def results(ctx, dA, xs, k, mode, indexed):
    ctx.set_option("packed_index", indexed)
    out = []
    y = ctx.empty(dA.n)
    dA.spmv(xs[0], y, mode=mode)
    out.append(y.to_host())
    kern = ctx.query("last_spmv_kernel")
    for v in dA.mpk(k, xs[0], mode=mode):
        out.append(v.to_host())
    strat = ctx.query("last_mpk_strategy")
    for row in dA.mpk_multi(k, xs, mode=mode):
        for v in row:
            out.append(v.to_host())
    return out, (kern, strat)


def timed(ctx, fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    return e0.elapsed_ms(e1) / reps


def main():
    ctx = nsk.Context(0)
    bad = 0
    ctx.set_option("wave_l2_pct", 1000)  # small operators: let them fuse
    modes = [nsk.EXACT_FMA, nsk.EXACT_MULADD] + ([nsk.FAST] if hasattr(nsk, "FAST") else [])
    for name, A in (("7pt 64x24x20", matgen.laplace3d_7pt(64, 24, 20)), ("7pt 61x17x23 (odd n)", matgen.laplace3d_7pt(61, 17, 23)),
                    ("5pt 300x41", matgen.laplace2d_5pt(300, 41)), ("7pt 128x64x48", matgen.laplace3d_7pt(128, 64, 48)),
                    ("random stencil", matgen.random_stencil3d(40, 24, 16, seed=1, max_points=6))):
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        xs = [ctx.to_device(matgen.vec_uniform(A.n, 3 + v)) for v in range(2)]
        for mode in modes:
            for k in (2, 4):
                ref, how0 = results(ctx, dA, xs, k, mode, 0)
                got, how1 = results(ctx, dA, xs, k, mode, 1)
                ok = all(same_bits(a, b) for a, b in zip(ref, got)) and how0 == how1
                bad += not ok
                if not ok:
                    print(f"MISMATCH {name} mode={mode} k={k} kernels {how0} / {how1}", flush=True)
        print(f"{name:24s} n={A.n:8d} kernels (spmv, mpk) {how0}: {'ok' if bad == 0 else 'FAILED so far: %d' % bad}", flush=True)
        dA.close()
    ctx.set_option("wave_l2_pct", 0)
    if bad:
        print(f"{bad} mismatches", flush=True)
        return 1

    A = matgen.laplace3d_7pt(256)
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    x2 = ctx.to_device(matgen.vec_uniform(A.n, 2))
    y = ctx.empty(A.n)
    print("256^3 ready", flush=True)
    res = {}
    for idx in (0, 1):
        ctx.set_option("packed_index", idx)
        ctx.axpy(1.0, x, y)
        dA.spmv(x, y)
        res[idx] = [y.to_host()]
        t_spmv = timed(ctx, lambda: dA.spmv(x, y))
        line = f"256^3 packed_index={idx}: SpMV {t_spmv:.4f} ms ({dA.spmv_bytes / t_spmv / 1e6:.0f} GB/s algorithmic)"
        for k in (2, 4):
            lv = [ctx.empty(A.n) for _ in range(k)]
            dA.mpk(k, x, lv)
            res[idx] += [v.to_host() for v in lv]
            t = timed(ctx, lambda: dA.mpk(k, x, lv))
            line += f" | k={k} {t:.4f} ms"
        lv2 = [[ctx.empty(A.n) for _ in range(4)] for _ in range(2)]
        t = timed(ctx, lambda: dA.mpk_multi(4, [x, x2], lv2), reps=10)
        line += f" | k=4 two vectors {t:.4f} ms"
        for l2 in (80, 90):
            ctx.set_option("wave_l2_pct", l2)
            lv = [ctx.empty(A.n) for _ in range(4)]
            t = timed(ctx, lambda: dA.mpk(4, x, lv), reps=10)
            line += f" | k=4 L2 {l2} % {t:.4f} ms"
        ctx.set_option("wave_l2_pct", 0)
        print(line, flush=True)
    ok = all(same_bits(a, b) for a, b in zip(res[0], res[1]))
    bad += not ok
    print("256^3 SpMV, k=2, k=4:", "bit-identical" if ok else "MISMATCH", flush=True)
    lv = [ctx.empty(A.n) for _ in range(4)]
    for idx in (0, 1):
        ctx.set_option("packed_index", idx)
        dA.mpk(4, x, lv)
        ctx.set_option("pk_timing", 1)
        print(f"# stage cycle, packed_index={idx}", file=sys.stderr, flush=True)
        dA.mpk(4, x, lv)
        ctx.sync()
        ctx.set_option("pk_timing", 0)
    ctx.set_option("packed_index", 0)
    print("index compression:", "all bit-identical" if bad == 0 else f"{bad} mismatches", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
