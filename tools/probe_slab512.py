"""Per-GPU slab of BASELINE config 5 on ONE GPU: 512 x 512 x 64 planes of the 7-point Laplacian (what each of 8 GPUs owns in
the 512^3 solve): powers k = 4 (the 512^2-row planes are too large a reach for a 4-level window: fused pairs), CG rates."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import navierstokes_b200 as nsk
from navierstokes_b200 import matgen
ctx = nsk.Context(0)
A = matgen.laplace3d_7pt(512, 512, 64)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
x = ctx.to_device(matgen.vec_uniform(A.n, 1))
def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = ctx.event(), ctx.event(); e0.record()
    for _ in range(reps): fn()
    e1.record(); return e0.elapsed_ms(e1) / reps
y = ctx.empty(A.n)
print(f"# 512x512x64 slab: n={A.n} nnz={A.nnz}")
ms = timed(lambda: dA.spmv(x, y)); print(f"spmv                      {ms:.4f} ms  {dA.spmv_bytes/ms/1e6:.0f} GB/s")
for k in (2, 4):
    lv = [ctx.empty(A.n) for _ in range(k)]
    l0 = ctx.launch_count; dA.mpk(k, x, lv); nl = ctx.launch_count - l0
    ms = timed(lambda: dA.mpk(k, x, lv))
    ctx.set_option("mpk_kernel", 1); ms1 = timed(lambda: dA.mpk(k, x, lv)); ctx.set_option("mpk_kernel", 0)
    print(f"powers k={k}: {nl} launch(es) {ms:.4f} ms ({k*dA.spmv_bytes/ms/1e6:.0f} GB/s SpMV-equivalent); as {k} products {ms1:.4f} ms")
xt = x; b = ctx.empty(A.n); dA.spmv(xt, b); xs = ctx.empty(A.n)
for s in (1, 4):
    dA.cg(b, xs, tol=1e-300, maxit=48, sstep=s); ctx.sync()
    t0 = time.perf_counter(); _, it, _, _ = dA.cg(b, xs, tol=1e-300, maxit=48, sstep=s); ctx.sync(); dt = time.perf_counter() - t0
    print(f"{'classical CG' if s == 1 else 's-step CG s=4'}: {it/dt:.1f} iterations/s ({dt/it*1e3:.3f} ms per iteration)")
