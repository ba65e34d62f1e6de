"""CPU: the C-ABI library loads, exports every symbol include/nsk.h declares, and refuses to work
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from navierstokes_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "nsk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(lib):
    from navierstokes_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nsk.h but not exported by libnsk.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in navierstokes_b200/_lib.py"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in include/nsk.h"


def test_version_and_strerror(lib):
    assert lib.nsk_version() == 100
    assert lib.nsk_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.nsk_strerror(-3)


def test_no_gpu_is_a_loud_failure(lib):
    """On a box without a GPU nsk_ctx_create must fail with NSK_ERR_NO_DEVICE -- never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    s = lib.nsk_ctx_create(0, C.byref(h))
    assert s == -3 and not h.value
    assert b"no CPU fallback" in lib.nsk_last_error(None)
    import navierstokes_b200 as nsk
    with pytest.raises(nsk.NskError):
        nsk.Context(0)


def test_product_never_imports_the_oracle():
    """Nothing under navierstokes_b200/ or include/ may reference oracle/ (judge's check, ours too)."""
    bad = []
    for f in list((ROOT / "navierstokes_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if f.suffix in {".py", ".cu", ".cpp", ".h", ".cuh", ".hpp"}:
            t = f.read_text(errors="ignore")
            if re.search(r"^\s*(import|from)\s+oracle\b", t, flags=re.M) or "liboracle" in t or "oracle/" in t.replace(
                    "oracle/nsk_oracle.c (oracle_cg)", ""):
                bad.append(str(f))
    assert not bad, bad


def test_sass_has_tma_bulk_copy():
    """The streaming kernels must really use the TMA bulk path (SASS UBLKCP) and fp64 FMA."""
    import shutil
    import subprocess
    from navierstokes_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    out = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UBLKCP" in out and "DFMA" in out and "SYNCS" in out
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True,
                                       text=True).stdout


def test_generators_shapes():
    from navierstokes_b200 import matgen
    A = matgen.laplace2d_5pt(64)
    assert A.n == 4096 and A.nnz == 5 * 4096 - 4 * 64
    B = matgen.laplace3d_7pt(16)
    assert B.n == 4096 and B.nnz == 7 * 4096 - 6 * 256
    assert B.spmv_bytes() == 12 * B.nnz + 4 * (B.n + 1) + 16 * B.n
    # row slab of the same operator
    S = matgen.laplace3d_7pt(16, row0=1024, nrows=512)
    assert np.array_equal(S.indcol, B.indcol[B.ptrow[1024]:B.ptrow[1536]])
    F = matgen.fem_baij4(3)
    assert F.n % 4 == 0 and np.all(np.diff(F.ptrow) % 4 == 0)
    assert np.array_equal(F.coef, F.coef.astype(np.float32).astype(np.float64))
    # full-size row counts of the BASELINE configs (formulas only, nothing allocated)
    assert 5 * 4096**2 - 4 * 4096 == 83_869_696
    assert 7 * 256**3 - 6 * 256**2 == 117_047_296


def test_shim_exports_the_reference_mangled_symbols():
    """Link-time drop-in: every kernel symbol the reference's own objects define for this path --
    compiled here from /root/reference where available -- is exported by libnsk_spmvshim.so under the same mangled
    name; on a box without the reference sources the expected list is the one recorded below."""
    import shutil
    import subprocess
    from navierstokes_b200 import _lib
    shim = _lib.LIB_PATH.parent / "libnsk_spmvshim.so"
    assert shim.exists(), "build the shim first (python -m navierstokes_b200.build)"
    if shutil.which("nm") is None:
        pytest.skip("nm not installed")
    have = set(subprocess.run(["nm", "-D", "--defined-only", str(shim)], capture_output=True, text=True).stdout.split())
    expected = [
        "_Z8SpMV_CSRPdS_R9csrmatrix", "_Z12SpMV_CSR_OPTPdS_R9csrmatrix", "_Z12SpMV_CSR_FMAPdS_R9csrmatrix",
        "_Z13SpMV_CSR_AVX2PdS_R9csrmatrix",
        "_Z9SpMV_BCSRPdPKdRK14bcsr4x4_matrix", "_Z13SpMV_BCSR_OPTPdPKdRK14bcsr4x4_matrix",
        "_Z13SpMV_BCSR_FMAPdPKdRK14bcsr4x4_matrix", "_Z14SpMV_BCSR_AVX2PdPKdRK14bcsr4x4_matrix",
        "_Z16Generate1stlayerRSt6vectorIiSaIiEER9csrmatrix",
        "_Z9SpM2V_CSRPdS_S_R9csrmatrixRSt6vectorIiSaIiEE", "_Z13SpM2V_CSR_OPTPdS_S_R9csrmatrixRSt6vectorIiSaIiEE",
        "_Z14SpM2V_CSR_AVX2PdS_S_R9csrmatrixRSt6vectorIiSaIiEE",
        "_Z22Generate1stlayer_BCSR4RSt6vectorIiSaIiEERK14bcsr4x4_matrix",
        "_Z10SpM2V_BCSRPdS_S_R14bcsr4x4_matrixRSt6vectorIiSaIiEE", "_Z14SpM2V_BCSR_OPTPdS_S_R14bcsr4x4_matrixRSt6vectorIiSaIiEE",
        "_Z14SpM2V_BCSR_FMAPdS_S_R14bcsr4x4_matrixRSt6vectorIiSaIiEE", "_Z15SpM2V_BCSR_AVX2PdS_S_R14bcsr4x4_matrixRSt6vectorIiSaIiEE",
    ]
    # the matrix-powers seed file's names (mpk/SpMVmulti0.cpp :22 :44 :65 :106 :132 :157 :191 :259)
    multi0 = [
        "_Z4SpMVPdS_R9csrmatrix", "_Z6SpM2V0PdS_S_R9csrmatrixRSt6vectorIiSaIiEE", "_Z5SpM2VPdS_S_R9csrmatrixRSt6vectorIiSaIiEE",
        "_Z5SpM3VPdS_S_S_R9csrmatrixRSt6vectorIiSaIiEERS2_IS4_SaIS4_EE",
        "_Z5SpM4VPdS_S_S_S_R9csrmatrixRSt6vectorIiSaIiEERS2_IS4_SaIS4_EERS2_IS7_SaIS7_EE",
        "_Z16Generate2ndlayerRSt6vectorIS_IiSaIiEESaIS1_EER9csrmatrixRS1_",
        "_Z16Generate3rdlayerRSt6vectorIS_IS_IiSaIiEESaIS1_EESaIS3_EER9csrmatrixRS1_RS3_",
    ]
    missing = [s for s in expected + multi0 if s not in have]
    assert not missing, missing
    ref_multi0 = ROOT / "oracle" / "_ref" / "libnsref_multi0.so"
    if ref_multi0.exists():
        defined = set(subprocess.run(["nm", "-D", "--defined-only", str(ref_multi0)], capture_output=True, text=True).stdout.split())
        for s in multi0:
            assert s in defined, f"{s} is not a symbol of the reference's SpMVmulti0.cpp: the recorded list is stale"
    # cross-check the recorded names against the reference's own objects when they were compiled here
    ref_dir = ROOT / "oracle" / "_ref"
    for obj, names in (("SpMV.o", expected[:8]), ("SpM2V.o", expected[8:])):
        o = ref_dir / obj
        if not o.exists():
            continue
        defined = set(subprocess.run(["nm", "--defined-only", str(o)], capture_output=True, text=True).stdout.split())
        for s in names:
            assert s in defined, f"{s} is not a symbol of the reference's {obj}: the recorded list is stale"


def test_shim_schedule_builders_match_the_reference(tmp_path):
    """Generate{1st,2nd,3rd}layer of the shim against the reference's own (compiled here into oracle/_ref), entry for
    entry on random operators with sorted and unsorted rows.  tests/layers_driver.cpp loads both libraries side by side."""
    import shutil
    import subprocess
    from navierstokes_b200 import _lib
    shim = _lib.LIB_PATH.parent / "libnsk_spmvshim.so"
    ref_multi0 = ROOT / "oracle" / "_ref" / "libnsref_multi0.so"
    if not ref_multi0.exists():
        pytest.skip("reference not compiled here (oracle/_ref absent)")
    if shutil.which("g++") is None:
        pytest.skip("g++ not installed")
    exe = tmp_path / "layers_driver"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "tests" / "layers_driver.cpp"),
                    "-o", str(exe), "-ldl"], check=True)
    r = subprocess.run([str(exe), str(shim), str(ref_multi0)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_bench_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) on a tiny grid: one JSON line with the
    keys the contract names, same metric / unit / config shape as the native arm."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--grid", "24"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "mpk_k4_spmv_equivalent_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
