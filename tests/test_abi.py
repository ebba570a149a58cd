"""CPU tier: the C-ABI library loads, exports every symbol include/stwo_b200.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            names |= set(re.findall(r"\b(stwo_b200_[a-z0-9_]+)\s*\(", open(os.path.join(inc, f)).read()))
    return names


def test_every_declared_symbol_is_exported_and_bound(pkg):
    lib = pkg._lib.load()
    declared = _declared()
    assert declared, "header declares nothing?"
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(pkg._lib.SIGNATURES), "python binding and header disagree: %s" % (declared ^ set(pkg._lib.SIGNATURES))


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = pkg._lib.load()
    assert lib.stwo_b200_init(0) == pkg._lib.E_NO_DEVICE
    st = np.zeros((2, 16), dtype=np.uint32)
    assert lib.stwo_b200_poseidon2_permute(st.ctypes.data_as(ctypes.c_void_p), 2) == pkg._lib.E_NO_DEVICE
    assert not st.any(), "a refused call must not touch the buffer"
    with pytest.raises(pkg.StwoB200Error):
        pkg.poseidon2_permute_host(st)


def test_product_never_touches_oracle():
    """Nothing under the product package may reference oracle/ (voids parity otherwise)."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "recursive-stwo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle|liborc|orc_", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
