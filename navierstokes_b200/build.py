"""Builds the sm_100a CUDA library in-tree: navierstokes_b200/lib/libnsk.so (+ the C++ shim).

nvcc cross-compiles without a GPU.  The built .so files are git-ignored but travel to the GPU box
with the repository snapshot, so they must live in-tree (not in a JIT cache).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _sources():
    return sorted(CSRC.glob("*.cu")) + [CSRC / "comm.cpp", CSRC / "dist_plan.cpp", CSRC / "ingest.cpp", CSRC / "reorder.cpp"]


def _stamp(files, extra=""):
    h = hashlib.sha256(extra.encode())
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "libnsk.so"
    deps = _sources() + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "nsk.h"]
    stamp = _stamp(deps, " ".join(NVCC_FLAGS))
    stamp_file = LIBDIR / "libnsk.stamp"
    if not force and out.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        build_shim(verbose=verbose, force=False)
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    jobs = jobs or min(8, os.cpu_count() or 1)
    procs = []
    objs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        while sum(p.poll() is None for _, p in procs) >= jobs:
            for _, p in procs:
                if p.poll() is None:
                    p.wait()
                    break
    for src, p in procs:
        outp, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{outp.decode()}")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(out), *map(str, objs), "-ldl"]
    subprocess.run(link, check=True)
    stamp_file.write_text(stamp)
    build_shim(verbose=verbose, force=True)
    return out


def build_shim(verbose: bool = False, force: bool = True) -> Path | None:
    """C++ drop-in library exporting the reference's SpMV.h symbols on top of libnsk.so."""
    src = CSRC / "shim_spmv.cpp"
    if not src.exists():
        return None
    out = LIBDIR / "libnsk_spmvshim.so"
    deps = [src, ROOT / "include" / "nsk_spmv_compat.hpp", ROOT / "include" / "nsk.h"]
    if not force and out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", str(ROOT / "include"), str(src), "-o", str(out),
           f"-L{LIBDIR}", "-lnsk", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
