"""Which SpMV kernel for operators with longer rows (C1 FEM-like 58/row, C4 tet 15/row)?  Exact mode, device-resident."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import navierstokes_b200 as nsk
from navierstokes_b200 import matgen

ctx = nsk.Context(0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = ctx.event(), ctx.event(); e0.record()
    for _ in range(reps): fn()
    e1.record(); return e0.elapsed_ms(e1) / reps

import sys as _s
which = _s.argv[1] if len(_s.argv) > 1 else "both"
cases = []
if which in ("both", "c1"):
    cases.append(("C1 fem_baij4(50)", matgen.fem_baij4(50)))
if which in ("both", "c4"):
    cases.append(("C4 tet(100) rcm", matgen.tet_p1_laplacian(100, permute_seed=2, rcm=True)))
for tag, A in cases:
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    x = ctx.to_device(matgen.vec_uniform(A.n, 1)); y = ctx.empty(A.n)
    B = dA.spmv_bytes
    print(f"# {tag}: n={A.n} nnz={A.nnz} mean {A.nnz/A.n:.1f} packed={dA.packed_bytes>0}", flush=True)
    ctx.set_option("spmv_kernel", 1); dA.spmv(x, y); ref = y.to_host()
    for kern, kind, var in [(1,0,0)] + [(3,0,v) for v in (0, 8, 11, 12, 13, 14)] + [(2,1,v) for v in (0,4)]:
        ctx.set_option("spmv_kernel", kern); ctx.set_option("stream_exact_kind", kind)
        ctx.set_option("packed_variant" if kern == 3 else "stream_variant", var)
        dA.spmv(x, y)
        ok = np.array_equal(y.to_host().view(np.int64), ref.view(np.int64))
        ms = timed(lambda: dA.spmv(x, y))
        print(f"kernel={kern} (ran {ctx.query('last_spmv_kernel')}) exact_kind={kind} variant={var}: {ms:.4f} ms {B/ms/1e6:8.1f} GB/s {'OK' if ok else 'MISMATCH'}", flush=True)
    ctx.set_option("spmv_kernel", 0); ctx.set_option("stream_exact_kind", 0); ctx.set_option("stream_variant", 0); ctx.set_option("packed_variant", 0)
    lv = [ctx.empty(A.n) for _ in range(2)]
    for strat in (0, 1):
        ctx.set_option("mpk_kernel", strat)
        ms = timed(lambda: dA.mpk(2, x, lv), 10)
        print(f"powers k=2 mpk_kernel={strat} (ran {ctx.query('last_mpk_strategy')}): {ms:.4f} ms  SpMV-equivalent {2*B/ms/1e6:8.1f} GB/s", flush=True)
    ctx.set_option("mpk_kernel", 0)
    ms = timed(lambda: dA.spmv(x, y, nsk.FAST))
    print(f"fast mode default: {ms:.4f} ms {B/ms/1e6:8.1f} GB/s", flush=True)
