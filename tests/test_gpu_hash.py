"""GPU tier: K1 / Merkle commit / K2 through the C ABI, bit-exact against the CPU oracle."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle_py as O

P = O.P
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _t(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)


def _n(t):
    return t.cpu().numpy().view(np.uint32)


def _edge_states(rng, n):
    st = rng.integers(0, P, size=(n, 16), dtype=np.uint32)
    st[0] = np.arange(16)
    st[1] = 0
    st[2] = P - 1
    st[3, :8] = P - 1
    st[3, 8:] = 0
    st[4, ::2] = P - 1
    return st


@pytest.mark.parametrize("variant", [0, 1])
def test_permute_kat_and_random(pkg, gpu, orc, rng, variant):
    kat = json.load(open(os.path.join(GOLD, "poseidon2_kat.json")))
    for n in (1, 5, 127, 128, 129, 10007):
        st = _edge_states(rng, max(n, 6))[:n] if n >= 6 else _edge_states(rng, 6)[:n]
        d = _t(st, gpu)
        pkg.poseidon2_permute(d, variant=variant)
        got = _n(d)
        assert np.array_equal(got, O.permute(st)), "n=%d variant=%d" % (n, variant)
        assert got[0].tolist() == kat["output"]


def test_permute_empty_and_host_entry(pkg, gpu, orc, rng):
    import torch
    pkg.poseidon2_permute(torch.empty((0, 16), dtype=torch.int32, device=gpu))
    st = _edge_states(rng, 3000)
    got = pkg.poseidon2_permute_host(st.copy())
    assert np.array_equal(got, O.permute(st))


def test_launch_counter_moves(pkg, gpu):
    import torch
    a = pkg.launch_count()
    pkg.poseidon2_permute(torch.zeros((4, 16), dtype=torch.int32, device=gpu))
    assert pkg.launch_count() == a + 1


@pytest.mark.parametrize("n_cols", [1, 4, 7, 8, 9, 16, 17, 50, 60])
def test_hash_node_batch(pkg, gpu, orc, rng, n_cols):
    n = 300
    cols = rng.integers(0, P, size=(n_cols, n), dtype=np.uint32)
    kids = rng.integers(0, P, size=(n, 16), dtype=np.uint32)
    leaf = _n(pkg.hash_node_batch(None, _t(cols, gpu), n))
    inner = _n(pkg.hash_node_batch(_t(kids, gpu), _t(cols, gpu), n))
    plain = _n(pkg.hash_node_batch(_t(kids, gpu), None, n))
    for i in (0, 1, 150, 299):
        assert (leaf[i] == O.hash_node(None, None, cols[:, i])).all()
        assert (inner[i] == O.hash_node(kids[i, :8], kids[i, 8:], cols[:, i])).all()
        assert (plain[i] == O.hash_node(kids[i, :8], kids[i, 8:], cols[:0, i])).all()


@pytest.mark.parametrize("log_n,n_cols,n_trees", [(0, 3, 2), (1, 8, 1), (5, 4, 3), (7, 9, 2), (8, 60, 2), (11, 8, 1), (12, 50, 1)])
def test_merkle_commit_matches_oracle(pkg, gpu, orc, rng, log_n, n_cols, n_trees):
    n = 1 << log_n
    cols = rng.integers(0, P, size=(n_trees, n_cols, n), dtype=np.uint32)
    nodes = _n(pkg.merkle_commit(_t(cols, gpu)))
    for t in range(n_trees):
        want = O.merkle_build(cols[t].T, log_n, n_cols)
        assert np.array_equal(nodes[t], want)
    roots = pkg.merkle_commit_host(cols)
    assert np.array_equal(roots, nodes[:, 0, :])


@pytest.mark.parametrize("log_n,n_cols,n_q", [(10, 4, 16), (10, 8, 32), (9, 50, 64), (8, 60, 128)])
def test_decommit_verify_roundtrip_and_tamper(pkg, gpu, orc, rng, log_n, n_cols, n_q):
    import torch
    n_trees = 3
    n = 1 << log_n
    cols = rng.integers(0, P, size=(n_trees, n_cols, n), dtype=np.uint32)
    index = rng.integers(0, n, size=(n_trees, n_q), dtype=np.uint32)
    index[0, 0], index[0, 1] = 0, n - 1
    d_cols, d_idx = _t(cols, gpu), _t(index, gpu)
    nodes = pkg.merkle_commit(d_cols)
    pcols, sib = pkg.merkle_decommit(d_cols, nodes, d_idx)
    roots = nodes[:, 0, :].contiguous()
    root_id = torch.arange(n_trees, dtype=torch.int32, device=gpu).repeat_interleave(n_q).contiguous()
    shape = pkg.PathShape.make(log_n, {log_n: n_cols})
    verdict, comp = pkg.merkle_path_verify(shape, d_idx.reshape(-1), pcols, sib, roots, root_id, want_roots=True)
    assert verdict.cpu().numpy().all()
    # oracle agrees on a sample of paths, from the raw gathered data
    h_cols, h_sib, h_comp = _n(pcols), _n(sib), _n(comp)
    for p in (0, 1, n_q, n_trees * n_q - 1):
        t, q = divmod(p, n_q)
        assert (h_cols[p] == cols[t, :, index[t, q]]).all()
        assert (O.path_root_mixed(log_n, {log_n: n_cols}, int(index[t, q]), h_cols[p], h_sib[p]) == h_comp[p]).all()
    # tamper: one flipped bit in a leaf value, a sibling, the index -> exactly those paths reject
    bad_cols, bad_sib, bad_idx = pcols.clone(), sib.clone(), d_idx.reshape(-1).clone()
    bad_cols[3, 0] ^= 1
    bad_sib[5, log_n - 1, 7] ^= 4
    bad_idx[7] ^= 1
    v = pkg.merkle_path_verify(shape, bad_idx, bad_cols, bad_sib, roots, root_id).cpu().numpy()
    assert sorted(np.nonzero(v == 0)[0].tolist()) == [3, 5, 7]
    # wrong root id
    wrong = root_id.clone()
    wrong[0] = 1
    v = pkg.merkle_path_verify(shape, d_idx.reshape(-1), pcols, sib, roots, wrong).cpu().numpy()
    assert v[0] == 0 and v[1:].all()
    # host-pointer entry point, same data
    v2, c2 = pkg.merkle_path_verify_host(shape, index.reshape(-1).copy(), h_cols.copy(), h_sib.copy(), _n(roots).copy(),
                                         _n(root_id).copy(), want_roots=True)
    assert v2.all() and np.array_equal(c2, h_comp)


def test_mixed_degree_paths(pkg, gpu, orc, rng):
    """columns injected at inner layers (trees 0-3 and the FRI first-layer tree of the fixtures)"""
    for depth, layers in ((15, {15: 4, 13: 4, 9: 4}), (13, {13: 50, 9: 10}), (13, {13: 60, 9: 12}), (1, {1: 8, 0: 3}), (0, {0: 9})):
        shape = pkg.PathShape.make(depth, layers)
        cpp, n_paths = shape.cols_per_path(), 37
        cols = rng.integers(0, P, size=(n_paths, cpp), dtype=np.uint32)
        sib = rng.integers(0, P, size=(n_paths, max(depth, 1), 8), dtype=np.uint32)[:, :depth].copy() if depth else np.zeros((n_paths, 0, 8), dtype=np.uint32)
        idx = rng.integers(0, 1 << depth, size=n_paths, dtype=np.uint32) if depth else np.zeros(n_paths, dtype=np.uint32)
        want = np.stack([O.path_root_mixed(depth, layers, int(idx[p]), cols[p], sib[p].reshape(-1, 8) if depth else np.zeros((1, 8), np.uint32)) for p in range(n_paths)])
        roots = want.copy()
        rid = np.arange(n_paths, dtype=np.uint32)
        roots[11, 0] ^= 1
        sib_dev = _t(sib, gpu) if depth else _t(np.zeros((1, 8), np.uint32), gpu)
        v, comp = pkg.merkle_path_verify(shape, _t(idx, gpu), _t(cols, gpu), sib_dev, _t(roots, gpu), _t(rid, gpu), want_roots=True)
        assert np.array_equal(_n(comp), want)
        v = v.cpu().numpy()
        assert v[11] == 0 and np.delete(v, 11).all()


def test_full_size_sweep_properties(pkg, gpu, orc):
    """BASELINE configs[2] at full size: 2^20-leaf tree, 128 queries.  Size-independent checks: every honest path
    accepts, the leaf layer and the root agree with the oracle's (sampled / recomputed from the GPU's own layers)."""
    import torch
    log_n, n_cols, n_q = 20, 8, 128
    n = 1 << log_n
    cols = O.synth_m31(7, n_cols * n).reshape(1, n_cols, n)
    idx = (O.splitmix64(7 ^ 0xABCDEF, n_q) & np.uint64(n - 1)).astype(np.uint32).reshape(1, n_q)
    d_cols, d_idx = _t(cols, gpu), _t(idx, gpu)
    nodes = pkg.merkle_commit(d_cols)
    pcols, sib = pkg.merkle_decommit(d_cols, nodes, d_idx)
    shape = pkg.PathShape.make(log_n, {log_n: n_cols})
    v = pkg.merkle_path_verify(shape, d_idx.reshape(-1), pcols, sib, nodes[:, 0, :].contiguous())
    assert v.cpu().numpy().all()
    h = _n(nodes)[0]
    for i in (0, 1, 12345, n - 1):                       # leaf hashes
        assert (h[n - 1 + i] == O.hash_node(None, None, cols[0, :, i])).all()
    for k in (19, 10, 3, 0):                             # inner layers: parent = hash_node(children)
        for i in (0, (1 << k) - 1):
            a, b = h[(2 << k) - 1 + 2 * i], h[(2 << k) - 1 + 2 * i + 1]
            assert (h[(1 << k) - 1 + i] == O.hash_node(a, b, cols[0, :0, 0])).all()
    # whole-tree checksum against the oracle built from the same leaves (2^21 permutations on the CPU: seconds)
    want = O.merkle_build(cols[0].T, log_n, n_cols)
    assert np.array_equal(h, want)
