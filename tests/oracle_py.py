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


# ---- full verifier -----------------------------------------------------------------------------------
MAX_INNER, MAX_Q, MAX_LOGS = 32, 128, 4
_Q = ctypes.c_uint32 * 4


class VerifyOut(ctypes.Structure):
    """orc_verify_out (oracle/orc.h)"""
    _fields_ = [(n, _Q) for n in ("z", "alpha", "random_coeff", "oods_t", "oods_x", "oods_y", "after_coeff")] + [
        ("fri_alphas", _Q * (MAX_INNER + 1)), ("digest_after_nonce", ctypes.c_uint32 * 8),
        ("raw_queries", ctypes.c_uint32 * MAX_Q), ("n_transcript_perms", ctypes.c_uint32),
        ("max_first_log", ctypes.c_uint32), ("n_inner", ctypes.c_uint32), ("n_queries", ctypes.c_uint32),
        ("n_logs", ctypes.c_uint32), ("log_sizes", ctypes.c_uint32 * MAX_LOGS),
        ("query_pos", (ctypes.c_uint32 * MAX_Q) * MAX_LOGS), ("oods_computed", _Q), ("oods_expected", _Q),
        ("domain_points", ((ctypes.c_uint32 * 2) * MAX_Q) * MAX_LOGS), ("fri_answers", (_Q * MAX_Q) * MAX_LOGS),
        ("circle_folds", (_Q * MAX_Q) * MAX_LOGS), ("line_folds", (_Q * MAX_Q) * MAX_INNER),
        ("last_layer_evals", _Q * MAX_Q), ("path_roots", ((ctypes.c_uint32 * 8) * MAX_Q) * (5 + MAX_INNER)),
        ("n_perms_hints", ctypes.c_uint64), ("n_perms_paths", ctypes.c_uint64), ("verdict", ctypes.c_int32),
        ("stage", ctypes.c_int32)]

    def arr(self, name):
        return np.ctypeslib.as_array(getattr(self, name)).copy()


STAGES = {0: "ok", 1: "parse", 2: "pow", 3: "logup", 4: "oods", 5: "merkle", 6: "fri_first", 7: "fri_inner", 8: "fri_last",
          9: "unsupported"}
INPUTS_SMALL = ([1], [[1, 0, 0, 0]])                                   # examples/single-proof/src/main.rs:28-33
INPUTS_RECURSIVE = ([1, 2, 3], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])   # examples/multi-proofs/src/main.rs (1,1),(2,i),(3,j)
PROOFS_DIR = os.path.join(ROOT, "tests", "golden", "proofs")


def inputs_for(name):
    return INPUTS_SMALL if name.startswith("small_proof") else INPUTS_RECURSIVE


def load_proof(name):
    """fixture -> 4-byte aligned uint8 array (padded), true length"""
    raw = np.fromfile(os.path.join(PROOFS_DIR, name), dtype=np.uint8)
    buf = np.zeros((raw.size + 3) // 4 * 4, dtype=np.uint8)
    buf[: raw.size] = raw
    return buf, raw.size


def verify_proof(buf, length, inputs, config=None):
    """config: (pow_bits, log_blowup, log_last, n_queries) the caller verifies under; None = whatever the proof's header says"""
    idx = np.array(inputs[0], dtype=np.uint32)
    vals = np.array(inputs[1], dtype=np.uint32)
    out = VerifyOut()
    if config is not None:
        cfg = np.array(config, dtype=np.uint32)
        load_oracle().orc_verify_proof_cfg(vp(buf), ctypes.c_size_t(length), vp(cfg), vp(idx), vp(vals), ctypes.c_uint32(idx.size), ctypes.byref(out))
        return out
    load_oracle().orc_verify_proof(vp(buf), ctypes.c_size_t(length), vp(idx), vp(vals), ctypes.c_uint32(idx.size), ctypes.byref(out))
    return out


def proof_offsets(buf, length):
    out = np.zeros(16, dtype=np.uint64)
    n = load_oracle().orc_proof_offsets(vp(buf), ctypes.c_size_t(length), vp(out))
    names = ["commitment0", "sampled0", "hash_witness0", "queried0", "fri_first_witness", "fri_first_hash_witness",
             "fri_first_commitment", "fri_inner0_witness", "fri_inner0_hash_witness", "last_coeffs", "queried3",
             "fri_inner_last_witness", "pow_nonce", "sampled3_7", "commitment3", "_"]
    return dict(zip(names, out.tolist())) if n else None
