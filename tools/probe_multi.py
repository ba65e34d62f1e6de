"""Two right-hand sides per launch vs one after the other (256^3, device-resident)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk
from navierstokes_b200 import matgen
ctx = nsk.Context(0)
A = matgen.laplace3d_7pt(256)
dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
xs = [ctx.to_device(matgen.vec_uniform(A.n, s)) for s in (1, 2)]
def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = ctx.event(), ctx.event(); e0.record()
    for _ in range(reps): fn()
    e1.record(); return e0.elapsed_ms(e1) / reps
for k in (1, 2, 4):
    lv = [[ctx.empty(A.n) for _ in range(k)] for _ in range(2)]
    t2 = timed(lambda: dA.mpk_multi(k, xs, lv))
    t1 = timed(lambda: (dA.mpk(k, xs[0], lv[0]), dA.mpk(k, xs[1], lv[1])))
    print(f"k={k}: two vectors fused {t2:.4f} ms, one after the other {t1:.4f} ms  ({2*k*dA.spmv_bytes/t2/1e6:.0f} vs {2*k*dA.spmv_bytes/t1/1e6:.0f} GB/s SpMV-equivalent)")
