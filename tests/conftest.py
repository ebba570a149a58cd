import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

P = (1 << 31) - 1


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _run(cmd, cwd=ROOT):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-2000:], r.stderr[-2000:]))


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liborc.so), built on demand."""
    from oracle_py import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def hostsim():
    """Device arithmetic headers compiled for the host (tests/hostsim)."""
    out = os.path.join(ROOT, "build", "libhostsim.so")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    csrc = os.path.join(ROOT, "recursive-stwo_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")] + [os.path.join(csrc, "dsl", f) for f in os.listdir(os.path.join(csrc, "dsl"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", out, src])
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes over libstwo_b200.so)."""
    return importlib.import_module("recursive-stwo_b200")


@pytest.fixture(scope="session")
def gpu(pkg):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    pkg.init(0)
    return torch.device("cuda:0")


@pytest.fixture
def rng():
    return np.random.default_rng(0xB200)


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)
