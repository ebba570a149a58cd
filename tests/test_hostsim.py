"""CPU tier: the kernels' own arithmetic headers (m31.cuh / poseidon2.cuh / merkle.cuh) compiled for the
host and compared with the oracle — catches lazy-reduction range bugs before any GPU time is spent."""
import ctypes
import importlib

import numpy as np

import oracle_py as O

P = O.P


def _edge_states(rng, n):
    st = rng.integers(0, P, size=(n, 16), dtype=np.uint32)
    st[0] = np.arange(16)
    st[1] = 0
    st[2] = P - 1
    st[3, :8] = P - 1
    st[3, 8:] = 0
    st[4, ::2] = P - 1
    st[5] = 1
    return st


def test_permute_both_shapes(hostsim, orc, rng):
    st = _edge_states(rng, 4096)
    want = O.permute(st)
    for fn in (hostsim.hs_poseidon2_permute, hostsim.hs_poseidon2_permute_rolled):
        got = st.copy()
        fn(O.vp(got), ctypes.c_size_t(got.shape[0]))
        assert np.array_equal(got, want)


def test_field_ops(hostsim, rng):
    hostsim.hs_m31_inv.restype = ctypes.c_uint32
    hostsim.hs_m31_mul.restype = ctypes.c_uint32
    hostsim.hs_m31_red64.restype = ctypes.c_uint32
    hostsim.hs_m31_red64.argtypes = [ctypes.c_uint64]
    vals = [1, 2, 3, P - 1, P - 2, 1 << 30, 12345678] + rng.integers(1, P, size=200).tolist()
    for a in vals:
        assert hostsim.hs_m31_inv(a) == pow(a, P - 2, P)
        for b in vals[:12]:
            assert hostsim.hs_m31_mul(a, b) == a * b % P
    for x in [0, 1, P, P + 1, 2 * P, (1 << 62) + 12345, (1 << 63) - 1, P * P, 2 * P * P] + rng.integers(0, 1 << 62, size=200).tolist():
        assert hostsim.hs_m31_red64(x) == x % P
    # QM31: (a + bu)(c + du), u^2 = 2 + i, i^2 = -1
    def cmul(x, y):
        return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
    def qmul(x, y):
        a, b, c, d = x[:2], x[2:], y[:2], y[2:]
        bd = cmul(b, d)
        r = ((2 * bd[0] - bd[1]) % P, (2 * bd[1] + bd[0]) % P)
        ac, ad, bc = cmul(a, c), cmul(a, d), cmul(b, c)
        return [(ac[0] + r[0]) % P, (ac[1] + r[1]) % P, (ad[0] + bc[0]) % P, (ad[1] + bc[1]) % P]
    for _ in range(200):
        x = rng.integers(0, P, size=4, dtype=np.uint32)
        y = rng.integers(0, P, size=4, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        hostsim.hs_qm31_mul(O.vp(x), O.vp(y), O.vp(o))
        assert o.tolist() == qmul(x.tolist(), y.tolist())
        hostsim.hs_qm31_inv(O.vp(x), O.vp(o))
        assert qmul(x.tolist(), o.tolist()) == [1, 0, 0, 0]


def test_circle_generator(hostsim):
    o = np.zeros(2, dtype=np.uint32)
    hostsim.hs_circle_mul_gen(ctypes.c_uint32(1), O.vp(o))
    assert o.tolist() == [2, 1268011823]
    assert (2 * 2 + 1268011823 ** 2) % P == 1
    hostsim.hs_circle_mul_gen(ctypes.c_uint32(1 << 30), O.vp(o))    # order-2 point
    assert o.tolist() == [P - 1, 0]


def test_hash_node_and_paths(hostsim, orc, rng):
    pkg = importlib.import_module("recursive-stwo_b200")
    for n in (1, 4, 7, 8, 9, 16, 17, 50, 60):
        cols = rng.integers(0, P, size=n, dtype=np.uint32)
        kids = rng.integers(0, P, size=16, dtype=np.uint32)
        out = np.zeros(8, dtype=np.uint32)
        hostsim.hs_hash_node(None, O.vp(cols), ctypes.c_uint32(n), O.vp(out))
        assert (out == O.hash_node(None, None, cols)).all()
        hostsim.hs_hash_node(O.vp(kids), O.vp(cols), ctypes.c_uint32(n), O.vp(out))
        assert (out == O.hash_node(kids[:8], kids[8:], cols)).all()
        hostsim.hs_hash_node(O.vp(kids), O.vp(cols), ctypes.c_uint32(0), O.vp(out))
        assert (out == O.hash_node(kids[:8], kids[8:], cols[:0])).all()
    # mixed-degree path shapes (FRI first layer tree of small_proof: 4 words at log 15, 13 and 9)
    for depth, layers in ((15, {15: 4, 13: 4, 9: 4}), (13, {13: 50, 9: 10}), (5, {5: 17}), (1, {1: 8, 0: 3}), (0, {0: 9})):
        shape = pkg.PathShape.make(depth, layers)
        cpp = shape.cols_per_path()
        for _ in range(4):
            cols = rng.integers(0, P, size=cpp, dtype=np.uint32)
            sib = rng.integers(0, P, size=(max(depth, 1), 8), dtype=np.uint32)
            idx = int(rng.integers(0, 1 << depth)) if depth else 0
            out = np.zeros(8, dtype=np.uint32)
            hostsim.hs_path_root(ctypes.byref(shape), ctypes.c_uint32(idx), O.vp(cols), O.vp(sib), O.vp(out))
            assert (out == O.path_root_mixed(depth, layers, idx, cols, sib)).all()
        assert pkg.path_perms(shape) == sum((n + 7) // 8 + 1 for n in layers.values()) + depth


def test_gate_check_split_over_two_lanes(hostsim, rng):
    """tape::gate_ok_half (the export kernel's fused check_arithmetics: one lane per CM31 half of the gate) == tape::gate_ok, on
    satisfied rows of every gate kind and on rows with one corrupted coordinate (constraint_system/src/plonk_with_poseidon.rs:337-380)"""
    import oracle_py as O
    P = O.P

    def qmul(x, y):
        out = np.zeros(4, dtype=np.uint32)
        hostsim.hs_qm31_mul(O.vp(x), O.vp(y), O.vp(out))
        return out
    n_bad = 0
    for k in range(600):
        a = rng.integers(0, P, 4, dtype=np.uint32)
        b = rng.integers(0, P, 4, dtype=np.uint32) if k % 5 else np.zeros(4, dtype=np.uint32)
        op = [0, 1, int(rng.integers(2, P))][k % 3]
        s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
        prod = qmul(a, b)
        c = ((s.astype(np.uint64) * op + prod.astype(np.uint64) * ((1 - op) % P)) % P).astype(np.uint32)
        enforce = 1 if (k % 7 == 0 and not c[1:].any()) else 0
        for corrupt in (None, k % 4):
            cc = c.copy()
            if corrupt is not None:
                cc[corrupt] = (int(cc[corrupt]) + 1 + k) % P
            whole = hostsim.hs_gate_ok(O.vp(a), O.vp(b), O.vp(cc), op, enforce)
            halves = hostsim.hs_gate_ok_half(O.vp(a), O.vp(b), O.vp(cc), op, enforce, 0) and hostsim.hs_gate_ok_half(O.vp(a), O.vp(b), O.vp(cc), op, enforce, 1)
            assert bool(whole) == bool(halves) == (corrupt is None), (k, corrupt)
            n_bad += corrupt is not None
        # enforce_c_m31 with a non-M31 c: both forms reject
        assert not hostsim.hs_gate_ok(O.vp(a), O.vp(b), O.vp(c), op, 1) or not c[1:].any()
        h = hostsim.hs_gate_ok_half(O.vp(a), O.vp(b), O.vp(c), op, 1, 0) and hostsim.hs_gate_ok_half(O.vp(a), O.vp(b), O.vp(c), op, 1, 1)
        assert bool(h) == bool(hostsim.hs_gate_ok(O.vp(a), O.vp(b), O.vp(c), op, 1))
    assert n_bad == 600
