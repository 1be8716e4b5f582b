import ctypes
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the -m gpu suites instead of failing them."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a B200 (run with -m gpu on the GPU box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _gxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU checkers (C restatement always; the reference build when oracle/_ref exists)."""
    import oracle
    oracle.load_port()
    return oracle


@pytest.fixture(scope="session")
def harness():
    """tests/host_harness.cpp: the product's lattice.cuh compiled for the host."""
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    hdr = os.path.join(ROOT, "tcam_wsol_video_b200", "csrc", "lattice.cuh")
    so = os.path.join(ROOT, "tests", "libhost_harness.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([_gxx(), "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src],
                       check=True)
    return ctypes.CDLL(so)


def golden_cases(prefix=""):
    return sorted(p for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")) if "224x224" not in p)


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_err(a, b):
    """Normwise relative error ||a-b||_inf / ||b||_inf."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def pointwise_rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())
