"""C++ drop-in shim on the GPU: a small driver written against the reference's names and STL containers is compiled
here, linked against libnsk_spmvshim.so and must reproduce a plain host loop bit for bit."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "navierstokes_b200" / "lib"


@pytest.mark.gpu
def test_cpp_shim_driver(tmp_path):
    assert shutil.which("g++"), "g++ is part of the image"
    exe = tmp_path / "shim_driver"
    cmd = ["g++", "-O2", "-std=c++11", "-ffp-contract=off", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "shim_driver.cpp"),
           "-o", str(exe), f"-L{LIB}", "-lnsk_spmvshim", "-lnsk", f"-Wl,-rpath,{LIB}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "shim_driver: OK" in r.stdout, r.stdout + r.stderr
