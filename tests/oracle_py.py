"""ctypes view of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
P = (1 << 31) - 1
_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liborc.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            r = subprocess.run(["make", "-C", ORACLE_DIR], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
        _lib = ctypes.CDLL(so)
        _lib.orc_merkle_build.restype = ctypes.c_uint64
    return _lib


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def permute(states):
    """states: [n,16] uint32 -> permuted copy"""
    s = np.ascontiguousarray(states, dtype=np.uint32).copy()
    load_oracle().orc_poseidon2_permute_batch(vp(s), ctypes.c_size_t(s.size // 16))
    return s


def hash_node(left, right, cols):
    out = np.zeros(8, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    l = vp(np.ascontiguousarray(left, dtype=np.uint32)) if left is not None else None
    r = vp(np.ascontiguousarray(right, dtype=np.uint32)) if right is not None else None
    load_oracle().orc_hash_node(l, r, vp(cols), ctypes.c_size_t(cols.size), vp(out))
    return out


def merkle_build(leaves_rowmajor, log_n, n_cols):
    """leaves_rowmajor: [2^log_n, n_cols] -> nodes [(2^(log_n+1)-1), 8] (root first)"""
    n = 1 << log_n
    nodes = np.zeros((2 * n - 1, 8), dtype=np.uint32)
    lv = np.ascontiguousarray(leaves_rowmajor, dtype=np.uint32)
    load_oracle().orc_merkle_build(vp(lv), ctypes.c_uint32(log_n), ctypes.c_uint32(n_cols), vp(nodes))
    return nodes


def path_root_mixed(depth, n_cols_by_h, index, cols, siblings):
    nc = np.zeros(depth + 1, dtype=np.uint32)
    for h, n in n_cols_by_h.items():
        nc[h] = n
    out = np.zeros(8, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    sib = np.ascontiguousarray(siblings, dtype=np.uint32)
    load_oracle().orc_merkle_path_root_mixed(ctypes.c_uint32(depth), vp(nc), ctypes.c_uint32(index), vp(cols), vp(sib), vp(out))
    return out


def splitmix64(seed, n):
    """n u64 outputs of splitmix64 seeded with `seed` (the synthetic-workload generator of BASELINE configs[2])."""
    out = np.empty(n, dtype=np.uint64)
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = x + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        out = z ^ (z >> np.uint64(31))
    return out


def synth_m31(seed, n):
    """n canonical M31 words from splitmix64(seed)"""
    return (splitmix64(seed, n) % np.uint64(P)).astype(np.uint32)
