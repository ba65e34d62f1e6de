import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def bits(a):
    """View float64 data as int64 so that comparisons are bit-exact (distinguishes -0.0, NaN payloads)."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} entries differ bitwise, first at {bad[:5]}"


def golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


CSR_CASES = ["lap3d_7pt_6", "lap3d_7pt_12x10x9", "lap2d_5pt_33x29", "fem_baij4_m3", "fem_baij4_m2_deep",
             "tet_p1_m5_rcm", "ragged_300"]
DEEP_CASES = ["lap3d_7pt_6", "fem_baij4_m2_deep", "tet_p1_m5_rcm"]
VECS = ["ones", "sin", "uni"]


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build(ref=False)
    return oracle.lib


@pytest.fixture(scope="session")
def ctx():
    """The GPU context.  No skip, no fallback: a missing library or GPU is a failure of a gpu test."""
    import navierstokes_b200 as nsk
    c = nsk.Context(0)
    yield c
    c.close()
