"""BASELINE config 4 at its stated size: unstructured tetrahedral P1 FEM Laplacian, ~50 M rows (368^3 nodes, ~15 entries
per row), randomly numbered then RCM-reordered with the library's own nsk_rcm, matrix powers k = 8 -- on one B200.

    python tools/bench_c4.py [--m 367] [--k 8] [--reps 5]

Times: the product (default kernel), k products one after the other, and the fused level pipeline (as many levels per
launch as the L2 window allows).  Parity: every level of the fused result against the k products bit for bit, and
8 random 65536-row slabs of the first level against the CPU oracle (SpMV_CSR_FMA restated, oracle/).
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import navierstokes_b200 as nsk  # noqa: E402
from navierstokes_b200 import _lib, matgen  # noqa: E402


def tetgen(m, perm_seed, threads):
    so = ROOT / "tools" / "bin" / "libtetgen.so"
    src = ROOT / "tools" / "gen" / "tetgen.cpp"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:  # built artefact (git-ignored): compile on first use
        import subprocess
        so.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", str(so), str(src)], check=True)
    lib = C.CDLL(str(so))
    lib.tetgen_rows.restype = C.c_int64
    lib.tetgen_rows.argtypes = [C.c_int, C.c_double, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n = (m + 1) ** 3
    ptrow = np.zeros(n + 1, np.int32)
    nnz = lib.tetgen_rows(m, 0.2, 1, perm_seed, ptrow.ctypes.data, None, None, threads)
    assert nnz > 0, "nnz does not fit the reference's int"
    indcol = np.empty(nnz, np.int32)
    coef = np.empty(nnz)
    lib.tetgen_rows(m, 0.2, 1, perm_seed, ptrow.ctypes.data, indcol.ctypes.data, coef.ctypes.data, threads)
    return matgen.Csr(n=n, ptrow=ptrow, indcol=indcol, coef=coef, ncols=n)


def timed(ctx, fn, reps, warm=2):
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
    ap.add_argument("--m", type=int, default=367)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=32)
    args = ap.parse_args()
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"config": f"C4 tet P1 Laplacian, Kuhn mesh {args.m + 1}^3 nodes, random numbering + RCM, k={args.k}"}
    t0 = time.time()
    A0 = tetgen(args.m, 2, args.threads)
    out["generate_s"] = round(time.time() - t0, 1)
    t0 = time.time()
    bw0 = matgen.bandwidth(A0)
    A = matgen.rcm_reorder(A0)
    del A0
    out["rcm_permute_s"] = round(time.time() - t0, 1)
    out.update(n=A.nrows, nnz=A.nnz, bandwidth_before=bw0, bandwidth_rcm=matgen.bandwidth(A))
    print(json.dumps(out), flush=True)
    ctx = nsk.Context(0)
    t0 = time.time()
    dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
    out["upload_s"] = round(time.time() - t0, 1)
    n, k = A.nrows, args.k
    x = matgen.vec_uniform(n, seed=1)
    dx = ctx.to_device(x)
    y = ctx.empty(n)
    B1, Bk = dA.spmv_bytes, dA.mpk_bytes(k)
    t0 = time.time()
    dA.spmv(dx, y)
    ctx.sync()
    out["first_product_s"] = round(time.time() - t0, 2)
    ms = timed(ctx, lambda: dA.spmv(dx, y), args.reps)
    out["spmv"] = {"ms": ms, "GBps": B1 / ms / 1e6, "frac_of_measured_peak": B1 / ms / 1e6 / peak,
                   "kernel": ctx.query("last_spmv_kernel")}
    # parity of the product against the oracle on random slabs
    import oracle
    oracle.build(ref=False)
    y1 = y.to_host()
    rng = np.random.default_rng(3)
    bad = 0
    for r0 in rng.integers(0, n - 65536, 8):
        r0 = int(r0)
        p0, p1 = int(A.ptrow[r0]), int(A.ptrow[r0 + 65536])
        ref = oracle.lib.spmv((A.ptrow[r0:r0 + 65537] - p0).astype(np.int32), A.indcol[p0:p1], A.coef[p0:p1], x)
        bad += int(np.count_nonzero(ref.view(np.int64) != y1[r0:r0 + 65536].view(np.int64)))
    out["spmv_oracle_slabs_entries_differing"] = bad
    lv = [ctx.empty(n) for _ in range(k)]
    ctx.set_option("mpk_kernel", 1)
    dA.mpk(k, dx, lv)
    ref = [l.to_host() for l in lv]
    ms1 = timed(ctx, lambda: dA.mpk(k, dx, lv), max(2, args.reps // 2))
    out["k_products"] = {"ms": ms1, "B_mpk_GBps": Bk / ms1 / 1e6}
    for name, opts in (("fused_auto", {"mpk_kernel": 0}), ("fused_sell", {"mpk_kernel": 5})):
        for o, v in opts.items():
            ctx.set_option(o, v)
        for l in lv:
            ctx.lib.nsk_memset0(ctx.h, l.ptr, 8 * n)
        t0 = time.time()
        l0 = ctx.launch_count
        dA.mpk(k, dx, lv)
        ctx.sync()
        first = time.time() - t0
        launches = ctx.launch_count - l0
        same = all(np.array_equal(lv[i].to_host().view(np.int64), ref[i].view(np.int64)) for i in range(k))
        ms = timed(ctx, lambda: dA.mpk(k, dx, lv), max(2, args.reps // 2))
        out[name] = {"ms": ms, "B_mpk_GBps": Bk / ms / 1e6, "frac_of_measured_peak": Bk / ms / 1e6 / peak,
                     "speedup_vs_k_products": ms1 / ms, "launches_per_call": launches, "strategy": ctx.query("last_mpk_strategy"),
                     "bitwise_equal_to_k_products": bool(same), "first_call_s": round(first, 2)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
