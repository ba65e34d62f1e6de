"""Secondary measurements on the other BASELINE.json configs (bench.py owns the headline one, C3).

    python tools/bench_configs.py [--quick] > profiles/rNN_configs.txt

C1  FEM-like BAIJ-4 operator (mmesh substitute): SpMV, CSR and 4x4 block CSR
C2  2D 5-point Poisson 4096^2: SpMV, matrix powers k = 2, 4
C3  3D 7-point Laplacian 256^3: classical CG and s-step CG (s = 4) iterations per second, solve to 1e-8
C4  P1 tet Laplacian, RCM-ordered (reduced size: the 50 M-row mesh takes minutes to assemble on the host): SpMV, k = 8
Device-resident data, CUDA events, 3 warm-ups; every operator is larger than L2 or the L2 is scrubbed between calls.
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

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timed(ctx, fn, reps, warm=3, flush=False):
    for _ in range(warm):
        fn()
    if not flush:
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        return e0.elapsed_ms(e1) / reps
    tot = 0.0
    for _ in range(reps):
        ctx.flush_l2()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        fn()
        e1.record()
        tot += e0.elapsed_ms(e1)
    return tot / reps


def report(name, ms, nbytes, extra=""):
    print(f"{name:58s} {ms:9.4f} ms  {nbytes/ms/1e6:9.1f} GB/s  {nbytes/ms/1e6/PEAK:6.3f} of measured peak {extra}", flush=True)


def spmv_and_mpk(ctx, tag, A, ks, reps, flush=False):
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1))
    y = ctx.empty(A.n)
    pb = dA.packed_bytes
    print(f"# {tag}: n={A.n} nnz={A.nnz} mean row {A.nnz/A.n:.1f}; packed format: "
          f"{'%.2f B/nnz' % (pb / A.nnz) if pb else 'not applicable (CSR kernels)'}", flush=True)
    ms = timed(ctx, lambda: dA.spmv(x, y), reps, flush=flush)
    report(f"{tag} SpMV exact (kernel {ctx.query('last_spmv_kernel')})", ms, dA.spmv_bytes)
    ms = timed(ctx, lambda: dA.spmv(x, y, nsk.FAST), reps, flush=flush)
    report(f"{tag} SpMV fast  (kernel {ctx.query('last_spmv_kernel')})", ms, dA.spmv_bytes)
    for k in ks:
        lv = [ctx.empty(A.n) for _ in range(k)]
        ms = timed(ctx, lambda: dA.mpk(k, x, lv), max(3, reps // 2), flush=flush)
        strat = ctx.query("last_mpk_strategy")
        report(f"{tag} powers k={k} (strategy {strat})", ms, dA.mpk_bytes(k), f" SpMV-equivalent {k*dA.spmv_bytes/ms/1e6:9.1f} GB/s")
        ctx.set_option("mpk_kernel", 1)
        ms1 = timed(ctx, lambda: dA.mpk(k, x, lv), max(3, reps // 2), flush=flush)
        ctx.set_option("mpk_kernel", 0)
        report(f"{tag} powers k={k} as {k} launches", ms1, dA.mpk_bytes(k), f" SpMV-equivalent {k*dA.spmv_bytes/ms1/1e6:9.1f} GB/s")
    return dA


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--c1", type=int, default=0, help="cells per edge of the C1 FEM-like operator (default 50 -> 530 k rows)")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    ctx = nsk.Context(0)
    reps = 5 if args.quick else 20
    print(f"# measured HBM copy peak {PEAK} GB/s; GB/s = algorithmic bytes (SURVEY.md 8d) / time", flush=True)

    if not only or "c2" in only:
        A = matgen.laplace2d_5pt(1024 if args.quick else 4096)
        spmv_and_mpk(ctx, "C2 2D 5-pt", A, (2, 4), reps)
        del A

    if not only or "c1" in only:
        t0 = time.time()
        A = matgen.fem_baij4(args.c1 if args.c1 > 0 else (24 if args.quick else 50))  # 51^3 nodes x 4 dof = 530 k rows, ~55 nnz/row (350 MB of CSR)
        print(f"# C1 assembled in {time.time()-t0:.1f}s", flush=True)
        dA = spmv_and_mpk(ctx, "C1 FEM BAIJ-4", A, (2,), reps)
        try:
            B = matgen.csr_to_bcsr4(A)
            dB = nsk.Bcsr4Matrix(ctx, B.ptrow, B.indcol, B.coef)
            xb = ctx.to_device(matgen.vec_uniform(A.n, 1))
            yb = ctx.empty(A.n)
            nblk = len(B.indcol)
            for batch in (1, 2, 4):
                ctx.set_option("bcsr_batch", batch)
                ms = timed(ctx, lambda: dB.spmv(xb, yb), reps)
                report(f"C1 FEM BAIJ-4 SpMV 4x4 block CSR, {batch} block(s) of loads in flight", ms,
                       128 * nblk + 4 * nblk + 4 * (A.n // 4 + 1) + 16 * A.n)
            ctx.set_option("bcsr_batch", 0)
        except Exception as e:  # generator helper may be absent
            print(f"# BCSR measurement skipped: {e}")
        del A, dA

    if not only or "c4" in only:
        t0 = time.time()
        m = 40 if args.quick else 100
        A = matgen.tet_p1_laplacian(m, permute_seed=2, rcm=True)
        print(f"# C4 (reduced: {m+1}^3 nodes instead of 368^3) assembled + RCM in {time.time()-t0:.1f}s", flush=True)
        spmv_and_mpk(ctx, "C4 tet P1 RCM", A, (8,), reps)
        del A

    if not only or "cg" in only:
        g = 96 if args.quick else 256
        A = matgen.laplace3d_7pt(g)
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        xt = ctx.to_device(matgen.vec_uniform(A.n, 1))
        b = ctx.empty(A.n)
        dA.spmv(xt, b)
        xs = ctx.empty(A.n)
        print(f"# C3/C5-per-GPU slab: 3D 7-pt {g}^3 (n={A.n}), b = A x_true, x0 = 0", flush=True)
        for s in (1, 4):
            # rate: a fixed number of iterations (tolerance unreachable), device-resident
            nit = 48
            dA.cg(b, xs, tol=1e-300, maxit=nit, sstep=s)
            ctx.sync()
            t0 = time.perf_counter()
            _, it, rel, ok = dA.cg(b, xs, tol=1e-300, maxit=nit, sstep=s)
            ctx.sync()
            dt = time.perf_counter() - t0
            print(f"{'classical CG' if s == 1 else 's-step CG s=%d' % s:20s}: {it} iterations in {dt*1e3:8.2f} ms -> {it/dt:8.1f} iterations/s "
                  f"({dt/it*1e3:.3f} ms/iteration)", flush=True)
        for s in (1, 4):
            ctx.sync()
            t0 = time.perf_counter()
            _, it, rel, ok = dA.cg(b, xs, tol=1e-8, maxit=5000, sstep=s)
            ctx.sync()
            dt = time.perf_counter() - t0
            err = float(np.max(np.abs(xs.to_host() - xt.to_host())))
            print(f"{'classical CG' if s == 1 else 's-step CG s=%d' % s:20s}: solve to 1e-8: {it} iterations, relres {rel:.2e}, converged={ok}, "
                  f"{dt*1e3:8.2f} ms, {it/dt:8.1f} iterations/s, max |x - x_true| = {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
