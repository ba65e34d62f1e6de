#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SpMV / matrix-powers path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the one the north-star target is quoted on): fp64 CSR 7-point Laplacian
256^3 (n = 16 777 216, nnz = 117 047 296), matrix powers k = 4.  One STEP = one call producing A x, A^2 x,
A^3 x, A^4 x.  N > 1: weak scaling by default -- every rank owns a 256 x 256 x 256 slab of a 256 x 256 x 256N
grid (row-partitioned operator, one depth-4 halo exchange per call over NCCL); `--scaling strong` splits the
256^3 problem instead.

Metric: SpMV-equivalent achieved GB/s = k * B_spmv / t with B_spmv = 12 nnz + 4(n+1) + 16 n, the
same definition for the GPU arm and the CPU reference arm.  `roofline` is the dominant kernel against the
measured HBM copy peak (MEASURED_PEAKS.json) using the COMPULSORY bytes of that kernel (fused powers kernel:
B_mpk = 12 nnz + 4(n+1) + 8n + 8nk per launch; plain SpMV kernel: B_spmv per launch).

The working set (1.74 GB per product) is far larger than the 126 MB L2, so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K_POWERS = 4
GRID = 256


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm: the UNMODIFIED reference kernel SpMV_CSR_FMA (oracle/_ref) run data-parallel over
# row slabs on every host core; falls back to the oracle port when _ref was not built.
# ---------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, A, threads: int):
        import oracle
        self.A = A
        self.threads = max(1, threads)
        self.kind = "reference" if oracle.ref.available() else "port"
        n = A.nrows
        bounds = np.linspace(0, n, self.threads + 1).astype(np.int64)
        self.slabs = []
        for t in range(self.threads):
            r0, r1 = int(bounds[t]), int(bounds[t + 1])
            p = (A.ptrow[r0:r1 + 1] - A.ptrow[r0]).astype(np.int32)
            c = A.indcol[A.ptrow[r0]:A.ptrow[r1]]
            v = A.coef[A.ptrow[r0]:A.ptrow[r1]]
            if self.kind == "reference":
                h = oracle.ref.csr(p, c, v)
                self.slabs.append((r0, r1, h))
            else:
                self.slabs.append((r0, r1, (np.ascontiguousarray(p), np.ascontiguousarray(c), np.ascontiguousarray(v))))
        self.oracle = oracle

    def spmv(self, x, y):
        def work(s):
            r0, r1, h = s
            if self.kind == "reference":
                h.spmv(x, "fma", out=y[r0:r1])
            else:
                p, c, v = h
                self.oracle.lib.l.oracle_spmv_csr_fma(r1 - r0, p, c, v, x, y[r0:r1])
        if self.threads == 1:
            work(self.slabs[0])
            return
        ts = [threading.Thread(target=work, args=(s,)) for s in self.slabs]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    def mpk(self, k, x, levels):
        src = x
        for l in range(k):
            self.spmv(src, levels[l])
            src = levels[l]


def slab_parity(matgen, nx, ny, nz_total, row_begin, n_owned, k, gpu_levels, threads):
    """Multi-GPU parity: this rank's owned rows of all k level vectors against k x SpMV_CSR_FMA on the CPU.

    The CPU multiplies the slab extended by k planes on either side (or up to the domain faces) as a truncated grid of
    its own: its rows are the global operator's rows except in the outermost plane of an artificial cut, and that error
    moves inwards one plane per level -- after k levels it has reached exactly the k extension planes, never an owned
    row.  The input is x[g] = sin(0.001 g) in GLOBAL numbering, so the ghost planes are part of the CPU input."""
    plane = nx * ny
    z0, z1 = row_begin // plane, (row_begin + n_owned) // plane
    e0, e1 = max(0, z0 - k), min(nz_total, z1 + k)
    Aext = matgen.laplace3d_7pt(nx, ny, e1 - e0)
    ref = CpuReference(Aext, threads)
    x = np.sin(0.001 * (e0 * plane + np.arange(Aext.nrows, dtype=np.float64)))
    lv = [np.zeros(Aext.nrows) for _ in range(k)]
    ref.mpk(k, x, lv)
    lo = (z0 - e0) * plane
    bad = 0
    for l in range(k):
        bad += int(np.count_nonzero(lv[l][lo:lo + n_owned].view(np.int64) != np.asarray(gpu_levels[l]).view(np.int64)))
    return bad, ref.kind


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, A, equiv_bytes):
    """--impl reference: the reference's CPU kernel on the host cores, same workload, same metric."""
    T = cpu_threads()
    ref = CpuReference(A, T)
    n = A.nrows
    x = np.sin(0.001 * np.arange(n))
    levels = [np.zeros(n) for _ in range(K_POWERS)]
    for _ in range(args.warmup):
        ref.mpk(K_POWERS, x, levels)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.mpk(K_POWERS, x, levels)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = equiv_bytes / dt / 1e9
    return {
        "impl": "reference", "metric": "mpk_k4_spmv_equivalent_GBps", "value": value, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": T, "kind": ref.kind,
                         "sample": f"{args.steps} full steps (k=4 x SpMV_CSR_FMA on the whole 256^3 operator), row slabs over "
                                   f"{T} host threads (the reference itself is single-threaded)"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def workload_config(args, world):
    return {"workload": f"3D 7-point Laplacian {GRID}^3 per GPU (fp64 CSR, int32 indices), matrix powers k={K_POWERS}"
                        if args.scaling == "weak" or world == 1 else
                        f"3D 7-point Laplacian {GRID}^3 split over {world} GPUs, matrix powers k={K_POWERS}",
            "n_rows_per_gpu": GRID ** 3 if args.scaling == "weak" or world == 1 else GRID ** 3 // world,
            "k": K_POWERS, "mode": "exact_fma (bit-identical to k x SpMV_CSR_FMA)",
            "partition": "1 GPU" if world == 1 else f"row slabs along z over {world} GPUs, depth-{K_POWERS} halo per call",
            "l2": "inputs larger than L2 (1.74 GB per product vs 126 MB), no flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--grid", type=int, default=GRID, help="grid edge (default 256; smaller only for debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the per-rank CPU parity check of the multi-GPU run")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record of a weak multi-GPU run")
    ap.add_argument("--c5", action="store_true", help="run the config-5 CG record (default: only with 8 GPUs)")
    ap.add_argument("--c5-grid", type=int, default=512, help="x/y extent of the config-5 grid; every GPU owns grid/8 planes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    globals()["GRID"] = args.grid

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from navierstokes_b200 import matgen

    if args.impl == "reference":
        if rank != 0:
            return 0
        A = matgen.laplace3d_7pt(GRID)
        print(json.dumps(run_reference_arm(args, A, K_POWERS * A.spmv_bytes())), flush=True)
        return 0

    import navierstokes_b200 as nsk
    peak, peak_src = measured_peaks()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        import datetime
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=90))
        dist = td

    ctx = nsk.Context(local_rank)
    n_local = GRID ** 3 if (args.scaling == "weak" or world == 1) else GRID ** 3 // world
    setup = None
    if world == 1:
        A = matgen.laplace3d_7pt(GRID)
        t0 = time.perf_counter()
        dA = nsk.CsrMatrix(ctx, A.ptrow, A.indcol, A.coef)
        ctx.sync()
        setup = {"operator_upload_ms": (time.perf_counter() - t0) * 1e3}
        spmv_bytes, mpk_bytes = dA.spmv_bytes, dA.mpk_bytes(K_POWERS)
        x_host = ctx.pinned(A.n)
        x_host[:] = np.sin(0.001 * np.arange(A.n))
        lv_host = [ctx.pinned(A.n) for _ in range(K_POWERS)]
        dx = ctx.to_device(x_host)
        dlv = [ctx.empty(A.n) for _ in range(K_POWERS)]
        step_dev = lambda: dA.mpk(K_POWERS, dx, dlv)           # noqa: E731
        step_e2e = lambda: dA.mpk(K_POWERS, x_host, lv_host)   # noqa: E731
    else:
        from navierstokes_b200 import distributed as nd
        nz_total = GRID * world if args.scaling == "weak" else GRID
        dop = nd.DistStencil3D(ctx, dist, GRID, GRID, nz_total, halo_depth=K_POWERS)
        A = dop.local_csr
        spmv_bytes, mpk_bytes = dop.spmv_bytes_owned, dop.mpk_bytes_owned(K_POWERS)
        x_host = ctx.pinned(dop.n_owned)
        x_host[:] = np.sin(0.001 * (dop.row_begin + np.arange(dop.n_owned)))
        lv_host = [ctx.pinned(dop.n_owned) for _ in range(K_POWERS)]
        dx = dop.new_vector(shared=True)  # registered: its halo is pushed over NVLink peer memory (no NCCL on the path)
        dop.set_owned(dx, x_host)
        dlv = [dop.new_vector() for _ in range(K_POWERS)]
        step_dev = lambda: dop.mpk(K_POWERS, dx, dlv)                     # noqa: E731
        step_e2e = lambda: dop.mpk_host(K_POWERS, x_host, lv_host)        # noqa: E731

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value, roofline) --------------------------------------------------------
    ctx.set_option("mpk_kernel", 0)  # default strategy: fused level pipeline (sliced-ELL pattern tiles for a stencil)
    t0 = time.perf_counter()
    step_dev()  # first call: builds the tile format of the operator (host packer) and the level schedule
    ctx.sync()
    if setup is not None:
        setup["first_call_ms"] = (time.perf_counter() - t0) * 1e3
        setup["resident_bytes"] = {"csr": int(12 * A.nnz + 4 * (A.n + 1)), "tiles": int(dA.tile_bytes)}
        setup["note"] = ("one-off per operator: upload of the caller's CSR arrays; first product = tile packing on the host "
                         "threads + upload + level schedule; both forms stay resident")
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    ms_total = e0.elapsed_ms(e1)
    launches = ctx.launch_count - launches0
    ms_step = max_over_ranks(ms_total / args.steps)
    # The timed region lasts a few milliseconds, nvidia-smi samples every 200 ms: keep the SAME load running (untimed)
    # for ~0.7 s so the sampler sees it.  The count is derived from the all-reduced step time, so every rank runs the
    # same number of extra steps (a distributed step contains a neighbour exchange: unmatched calls would deadlock).
    extra = int(min(3000, max(20, 700.0 / max(ms_step, 1e-3))))
    for _ in range(extra):
        step_dev()
    barrier()
    clocks = sampler.stop()
    equiv_total = K_POWERS * spmv_bytes * world
    value = equiv_total / ms_step / 1e6

    # dominant kernel: launches per step tell which strategy ran
    per_step = launches / max(args.steps, 1)
    fused = per_step < K_POWERS
    strategy = ctx.query("last_mpk_strategy")
    fused_names = {4: "packed_kernel (level pipeline over the packed format)",
                   5: "sell_tma_kernel (level pipeline over sliced-ELL pattern tiles, coefficients staged by bulk copies)"}
    if fused:
        kern_bytes, kern_ms = mpk_bytes, ms_total / args.steps
        kern_name = f"{fused_names.get(strategy, 'fused powers kernel')}, 1 launch per step"
    else:
        kern_bytes, kern_ms = spmv_bytes, ms_total / args.steps / K_POWERS
        kern_name = ({3: "packed_kernel", 4: "sell_tma_kernel"}.get(ctx.query("last_spmv_kernel"), "spmv_stream_kernel")
                     + f", {K_POWERS} launches per step")
    achieved = kern_bytes / kern_ms / 1e6
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, from the committed ncu capture
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists() and world == 1:
        try:
            t = json.loads(tfile.read_text())
            key = "fused" if fused else "spmv"
            if t.get(key, {}).get("strategy") == (strategy if fused else ctx.query("last_spmv_kernel")):
                traffic = t[key]["dram_bytes_per_launch"]
        except Exception:
            traffic = None

    # secondary: plain SpMV rate of the same operator (one product per launch)
    y = dlv[0]
    if world == 1:
        spmv_once = lambda: dA.spmv(dx, y)   # noqa: E731
    else:
        spmv_once = lambda: dop.spmv(dx, y)  # noqa: E731
    for _ in range(3):
        spmv_once()
    s0, s1 = ctx.event(), ctx.event()
    s0.record()
    for _ in range(max(args.steps, 10)):
        spmv_once()
    s1.record()
    spmv_ms = max_over_ranks(s0.elapsed_ms(s1) / max(args.steps, 10))
    spmv_traffic = None
    if tfile.exists() and world == 1:
        try:
            t = json.loads(tfile.read_text())
            if t.get("spmv", {}).get("strategy") == ctx.query("last_spmv_kernel"):
                spmv_traffic = t["spmv"]["dram_bytes_per_launch"]
        except Exception:
            spmv_traffic = None

    # ---- end to end through the public host-pointer API (pinned host buffers, copies inside) -------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    checksum = float(lv_host[K_POWERS - 1][:1024].sum())  # the D2H result is really read

    out = {
        "metric": "mpk_k4_spmv_equivalent_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": equiv_total / e2e_ms / 1e6, "unit": "GB/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 8 * n_local * world, "d2h_bytes_per_step": 8 * n_local * K_POWERS * world,
                "note": "nsk_mpk with NSK_HOST pointers: pinned x in, k level vectors out, operator resident in HBM; PCIe-bound "
                        "(the host-pointer call runs the k products one launch each and copies level l out on a second stream "
                        "while level l+1 is computed, so the fused kernel is not on this path; device-resident callers such "
                        "as nsk_cg get `value`)",
                "checksum": checksum},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kern_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(kern_bytes)},
        "spmv": {"ms": spmv_ms, "achieved_GBps": spmv_bytes * world / spmv_ms / 1e6,
                 "frac_of_peak": spmv_bytes / spmv_ms / 1e6 / peak, "algorithmic_bytes": int(spmv_bytes),
                 "traffic": spmv_traffic,
                 "traffic_GBps": (spmv_traffic / spmv_ms / 1e6) if spmv_traffic else None,
                 "note": "frac_of_peak divides CSR bytes (12 B/nnz + vectors) by the time of a kernel that streams an "
                         "index-free tile format (8 B/nnz + 1 B/row): above 1 by construction; traffic_GBps = ncu DRAM "
                         "bytes of that kernel / its time is the physical rate"},
        "mpk_bytes_rate_GBps": mpk_bytes * world / ms_step / 1e6,
        "value_note": "value = k * B_spmv / t (what k separate products would have to move; above the HBM peak by construction "
                      "when the fused kernel reads the operator once). The compulsory-bytes rate of the fused kernel is "
                      "mpk_bytes_rate_GBps = roofline.achieved (B_mpk / t), the number to hold against the HBM peak.",
    }
    if setup is not None:
        out["setup"] = setup
    # ---- multi-GPU parity: every rank checks its own slab of all k levels against the CPU reference ---------------
    if world > 1 and not args.no_parity:
        import torch
        bad, kind = slab_parity(matgen, GRID, GRID, nz_total, dop.row_begin, dop.n_owned, K_POWERS, lv_host,
                                max(1, cpu_threads() // world))
        t = torch.tensor([1.0 if bad == 0 else 0.0, float(bad)], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out["parity"] = {"ranks_ok": int(t[0].item()), "ranks": world, "bitwise": bool(int(t[0].item()) == world),
                         "entries_differing": int(t[1].item()), "against": f"k x SpMV_CSR_FMA ({kind}) on each rank's slab "
                         f"extended by {K_POWERS} ghost planes, all {K_POWERS} level vectors, owned rows"}
    # ---- strong scaling beside the weak number: the SAME 256^3 problem split over the ranks ---------------------
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        try:
            import torch
            from navierstokes_b200 import distributed as nd
            sop = nd.DistStencil3D(ctx, dist, GRID, GRID, GRID, halo_depth=K_POWERS)
            sx_host = np.sin(0.001 * (sop.row_begin + np.arange(sop.n_owned)))
            sdx = sop.new_vector(shared=True)
            sop.set_owned(sdx, sx_host)
            sdlv = [sop.new_vector() for _ in range(K_POWERS)]
            for _ in range(5):
                sop.mpk(K_POWERS, sdx, sdlv)
            barrier()
            g0, g1 = ctx.event(), ctx.event()
            g0.record()
            for _ in range(args.steps):
                sop.mpk(K_POWERS, sdx, sdlv)
            g1.record()
            s_ms = max_over_ranks(g0.elapsed_ms(g1) / args.steps)
            # halo exchange alone (depth k), same operator
            for _ in range(3):
                sop.halo_exchange(sdx, K_POWERS)
            barrier()
            g0.record()
            for _ in range(args.steps):
                sop.halo_exchange(sdx, K_POWERS)
            g1.record()
            h_ms = max_over_ranks(g0.elapsed_ms(g1) / args.steps)
            ctx.set_option("halo_push", 0)  # the same exchange through pack + ncclSend/ncclRecv + unpack, for comparison
            for _ in range(3):
                sop.halo_exchange(sdx, K_POWERS)
            barrier()
            g0.record()
            for _ in range(args.steps):
                sop.halo_exchange(sdx, K_POWERS)
            g1.record()
            hn_ms = max_over_ranks(g0.elapsed_ms(g1) / args.steps)
            ctx.set_option("halo_push", 1)
            sbad = 0
            if not args.no_parity:
                got = [sop.get_owned(v) for v in sdlv]
                sbad, _ = slab_parity(matgen, GRID, GRID, GRID, sop.row_begin, sop.n_owned, K_POWERS, got,
                                      max(1, cpu_threads() // world))
            t = torch.tensor([1.0 if sbad == 0 else 0.0], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            n_tot = GRID ** 3
            nnz_tot = 7 * n_tot - 6 * GRID * GRID
            spmv_tot = 12 * nnz_tot + 4 * (n_tot + 1) + 16 * n_tot
            out["strong"] = {"workload": f"3D 7-point Laplacian {GRID}^3 split over {world} GPUs (z-slabs), k={K_POWERS}",
                             "ms_per_step": s_ms, "value": K_POWERS * spmv_tot / s_ms / 1e6, "unit": "GB/s",
                             "halo_exchange_us": h_ms * 1e3, "halo_exchange_nccl_us": hn_ms * 1e3,
                             "efficiency_vs_n1_note": "t(1 GPU) / (N * t(N GPUs)) with t(1 GPU) = this build's N=1 ms_per_step "
                                                      "(the driver's SCALE record); not computed here",
                             "parity": {"ranks_ok": int(t[0].item()), "ranks": world, "bitwise": bool(int(t[0].item()) == world)}}
            sop.close()
        except Exception as exc:  # a failing sub-record must not take the headline line with it
            out["strong"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    # ---- BASELINE config 5 on the full box: pressure-Poisson CG on 512^3 over 8 GPUs, classical and s-step (s = 4) ----
    if world > 1 and not args.no_cg and (world == 8 or args.c5):
        try:
            import torch
            from navierstokes_b200 import distributed as nd
            del dlv, dx
            nxy, planes = args.c5_grid, args.c5_grid // 8
            t0 = time.perf_counter()
            cop = nd.DistStencil3D(ctx, dist, nxy, nxy, planes * world, halo_depth=K_POWERS)
            plan_s = time.perf_counter() - t0
            gidx = cop.row_begin + np.arange(cop.n_owned)
            x_true = np.sin(0.001 * gidx) + 0.5
            cdx, cdb = cop.new_vector(), cop.new_vector()
            cop.set_owned(cdx, x_true)
            cop.spmv(cdx, cdb)
            b_own = cop.get_owned(cdb)
            rec = {"workload": f"7-point Laplacian {nxy}x{nxy}x{planes * world} over {world} GPUs (z-slabs), b = A x_true, x0 = 0, "
                               f"to ||r||/||b|| <= 1e-8", "plan_s": plan_s}
            for s_step, name in ((1, "classical"), (4, "sstep4")):
                cop.cg(b_own, tol=1e-300, maxit=8, sstep=s_step)  # warm: plans, tile format, workspaces
                barrier()
                t0 = time.perf_counter()
                xs, it, rel, ok = cop.cg(b_own, tol=1e-8, maxit=4000, sstep=s_step)
                barrier()
                dt = time.perf_counter() - t0
                # true residual, recomputed from the returned solution with one more distributed product
                cop.set_owned(cdx, xs)
                cop.spmv(cdx, cdb)
                r = b_own - cop.get_owned(cdb)
                t = torch.tensor([float(r @ r), float(b_own @ b_own), 0.0], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                e = torch.tensor([float(np.max(np.abs(xs - x_true)))], dtype=torch.float64, device=f"cuda:{local_rank}")
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
                rec[name] = {"iters_to_1e-8": int(it), "converged": bool(ok), "iters_per_s": it / dt, "solve_ms": dt * 1e3,
                             "relres_recurrence": float(rel), "true_relres": float((t[0] / t[1]).sqrt().item()),
                             "max_abs_err_vs_x_true": float(e.item())}
            rec["note"] = ("host wall clock around nsk_cg with the owned parts of b and x in host memory (copies included); parity of CG "
                           "is unpinned (the reference has no CG): checked through the true residual")
            out["cg"] = rec
            cop.close()
        except Exception as exc:  # a failing sub-record must not take the headline line with it
            out["cg"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    # ---- Poisson CG iterations per second on the same operator (BASELINE metric, second half) ------------------
    if world == 1 and not args.no_cg:
        try:
            b = ctx.empty(A.n)
            dA.spmv(dx, b)
            xs = ctx.empty(A.n)
            cg = {}
            for s_step, name in ((1, "classical"), (4, "sstep4")):
                nit = 48
                dA.cg(b, xs, tol=1e-300, maxit=nit, sstep=s_step)  # warm: plans, workspace
                ctx.sync()
                t0 = time.perf_counter()
                _, it, _, _ = dA.cg(b, xs, tol=1e-300, maxit=nit, sstep=s_step)
                ctx.sync()
                cg[name + "_iters_per_s"] = it / (time.perf_counter() - t0)
            cg["note"] = ("48 iterations each on the 256^3 operator, device-resident b and x, host wall clock around the whole "
                          "nsk_cg call (includes its workspace allocation)")
            out["cg"] = cg
            del b, xs
        except Exception as exc:  # a failing sub-record must not take the headline line with it
            out["cg"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        T = cpu_threads()
        ref = CpuReference(A, T)
        xs = np.asarray(x_host).copy()
        lv = [np.zeros(A.nrows) for _ in range(K_POWERS)]
        ref.mpk(K_POWERS, xs, lv)  # warm
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ref.mpk(K_POWERS, xs, lv)
        dt = (time.perf_counter() - t0) / reps
        # the CPU result doubles as a full-size parity check of the GPU path (bit-exact mode)
        same = all(np.array_equal(lv[l].view(np.int64), np.asarray(lv_host[l]).view(np.int64)) for l in range(K_POWERS))
        ref1 = CpuReference(A, 1)
        t1 = time.perf_counter()
        ref1.spmv(xs, lv[0])
        dt1 = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": K_POWERS * spmv_bytes / dt / 1e9, "unit": "GB/s", "cores": T, "kind": ref.kind,
                               "sample": f"{reps} full steps of the same workload (k=4 x SpMV_CSR_FMA, whole 256^3 operator), "
                                         f"row slabs over {T} host threads (the reference itself is single-threaded: "
                                         f"single_thread_spmv_GBps)",
                               "single_thread_spmv_GBps": spmv_bytes / dt1 / 1e9,
                               "gpu_bitwise_equal_to_cpu_reference": bool(same)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
