"""Shared by the circuit parity tests (host-sim and GPU): oracle circuit per fixture and the wiring comparison."""
import os
import sys

import numpy as np

import oracle_py as O

sys.path.insert(0, O.ORACLE_DIR)
import orc_dsl as D  # noqa: E402

_cache = {}


def oracle_circuit(name, multipliers=1):
    """(CS, VerifyOut) of the oracle's circuit DSL for a fixture (cached per session)."""
    key = (name, multipliers)
    if key not in _cache:
        blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
        inputs = D.INPUTS_SINGLE if name.startswith("small") else D.INPUTS_RECURSIVE
        _cache[key] = D.verifier_circuit(blob, inputs, multipliers, O.VerifyOut)
    return _cache[key]


_cache_last = {}


def oracle_last_circuit(name):
    """(CSWithout, VerifyOut) of the oracle's last-layer circuit for a Poseidon31 fixture"""
    if name not in _cache_last:
        blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
        _cache_last[name] = D.last_layer_circuit(blob, O.VerifyOut)
    return _cache_last[name]


def compare_wiring_without(cs, info, get):
    """the Plonk-without-Poseidon flavour of compare_wiring: a/b/c wires and the four selectors"""
    n_rows = len(cs.a_wire)
    assert (info["n_rows"], info["n_rows_unpadded"], info["n_vars"], info["num_input"], info["n_flow"]) == (
        n_rows, cs.n_rows_unpadded, len(cs.variables), cs.num_input, 0)
    follows = get(6, n_rows).astype(bool)
    for what, name in ((0, "a_wire"), (1, "b_wire"), (2, "c_wire"), (5, "op1"), (11, "op2"), (12, "op3"), (13, "op4")):
        got, want = get(what, n_rows), np.array(getattr(cs, name), dtype=np.uint32)
        if name == "op1":
            c_val = np.array(cs.variables, dtype=np.uint32)[np.array(cs.c_wire)][:, 0]
            assert np.array_equal(want[follows], c_val[follows])
            got, want = got[~follows], want[~follows]
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, "%s differs first at row %d: %d != %d" % (name, bad[0], got[bad[0]], want[bad[0]])
    return follows


WIRING = ("a_wire", "b_wire", "c_wire", "poseidon_wire", "enforce_c_m31", "op")


def compare_wiring(cs, info, get):
    """cs: oracle CS; info: dict from the product recorder; get(what, n) -> uint32 array of a recorded column.
    Returns the op_follows_c mask."""
    n_rows = len(cs.a_wire)
    assert (info["n_rows"], info["n_rows_unpadded"], info["n_vars"]) == (n_rows, cs.n_rows_unpadded, len(cs.variables))
    assert (info["n_flow"], info["n_flow_padded"], info["num_input"]) == (cs.n_flow_unpadded, cs.n_flow_padded, cs.num_input)
    follows = get(6, n_rows).astype(bool)
    for what, name in enumerate(WIRING):
        got, want = get(what, n_rows), np.array(getattr(cs, name), dtype=np.uint32)
        if name == "op":
            # rows whose op constant follows the selected value (circle/src/lib.rs:80-92) are per-proof: c = op there
            c_val = np.array(cs.variables, dtype=np.uint32)[np.array(cs.c_wire)][:, 0]
            assert np.array_equal(want[follows], c_val[follows])
            got, want = got[~follows], want[~follows]
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, "%s differs first at row %d: %d != %d" % (name, bad[0], got[bad[0]], want[bad[0]])
    wire, addr, h, sw = cs.flow_arrays()
    assert np.array_equal(get(7, wire.size).reshape(-1, 4), wire)
    assert np.array_equal(get(8, addr.size), addr)
    return follows
