"""CPU tier: the oracle against the reference's golden vectors and its own invariants."""
import ctypes
import json
import os

import numpy as np

import oracle_py as O

P = O.P
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_poseidon2_kat(orc):
    kat = json.load(open(os.path.join(GOLD, "poseidon2_kat.json")))
    out = O.permute(np.array([kat["input"]], dtype=np.uint32))
    assert out[0].tolist() == kat["output"]


def test_permute_is_a_permutation_of_canonical_words(orc, rng):
    st = rng.integers(0, P, size=(512, 16), dtype=np.uint32)
    out = O.permute(st)
    assert (out < P).all()
    assert len({tuple(r) for r in out.tolist()}) == 512


def test_hash_node_structure(orc, rng):
    """hash_node follows primitives/merkle/src/lib.rs: sponge of 8-word chunks, then finalise / combine."""
    for n in (1, 7, 8, 9, 13, 16, 17, 21, 25, 60):
        cols = rng.integers(0, P, size=n, dtype=np.uint32)
        cap = np.zeros(8, dtype=np.uint32)
        for c in range((n + 7) // 8):
            chunk = np.zeros(8, dtype=np.uint32)
            part = cols[8 * c: 8 * c + 8]
            chunk[: part.size] = part
            cap = O.permute(np.concatenate([chunk, cap])[None])[0][8:]
        leaf = O.permute(np.concatenate([np.zeros(8, dtype=np.uint32), cap])[None])[0][:8]
        assert (O.hash_node(None, None, cols) == leaf).all()
        l, r = rng.integers(0, P, size=(2, 8), dtype=np.uint32)
        tree = O.permute(np.concatenate([l, r])[None])[0][:8]
        assert (O.hash_node(l, r, np.zeros(0, dtype=np.uint32)) == tree).all()
        comb = O.permute(np.concatenate([tree, cap])[None])[0][:8]
        assert (O.hash_node(l, r, cols) == comb).all()


def test_merkle_build_and_paths(orc, rng):
    log_n, n_cols = 6, 10
    leaves = rng.integers(0, P, size=(1 << log_n, n_cols), dtype=np.uint32)
    nodes = O.merkle_build(leaves, log_n, n_cols)
    root = nodes[0]
    for idx in (0, 1, 37, 63):
        sib = np.stack([nodes[(1 << (log_n - l)) - 1 + ((idx >> l) ^ 1)] for l in range(log_n)])
        got = O.path_root_mixed(log_n, {log_n: n_cols}, idx, leaves[idx], sib)
        assert (got == root).all()
        bad = sib.copy()
        bad[2, 3] ^= 1
        assert not (O.path_root_mixed(log_n, {log_n: n_cols}, idx, leaves[idx], bad) == root).all()


def test_channel(orc):
    """primitives/channel/src/lib.rs:23-58: mix -> digest = capacity, draw -> rate with counter, digest unchanged."""
    class Ch(ctypes.Structure):
        _fields_ = [("digest", ctypes.c_uint32 * 8), ("n_sent", ctypes.c_uint32), ("n_perms", ctypes.c_uint64)]
    c = Ch()
    orc.orc_channel_init(ctypes.byref(c))
    root = np.arange(1, 9, dtype=np.uint32)
    orc.orc_channel_mix_root(ctypes.byref(c), O.vp(root))
    exp = O.permute(np.concatenate([root, np.zeros(8, dtype=np.uint32)])[None])[0]
    assert list(c.digest) == exp[8:].tolist() and c.n_sent == 0
    out = np.zeros(8, dtype=np.uint32)
    orc.orc_channel_draw(ctypes.byref(c), O.vp(out))
    st = np.zeros(16, dtype=np.uint32)
    st[8:] = exp[8:]
    assert (out == O.permute(st[None])[0][:8]).all() and c.n_sent == 1
    orc.orc_channel_draw(ctypes.byref(c), O.vp(out))
    st[0] = 1
    assert (out == O.permute(st[None])[0][:8]).all() and list(c.digest) == exp[8:].tolist()
