"""CPU tier: the product's circuit recorder (csrc/dsl/*.hpp, host C++) and tape evaluator (csrc/tape.cuh, compiled for the
host) against the oracle's circuit DSL: identical wiring row by row, identical variables[], Poseidon flow and verdict of
check_arithmetics for the reference's fixtures -- the same code the GPU runs one lane per proof."""
import ctypes

import numpy as np
import pytest

import oracle_py as O
from circuit_common import compare_wiring, compare_wiring_without, oracle_circuit, oracle_last_circuit
from verify_common import Detail, pack, shape_of

INFO = ("n_rows", "n_rows_unpadded", "n_vars", "n_flow", "n_flow_padded", "n_input_words", "n_ins", "n_levels", "num_input",
        "words_per_instance")


def record(hs, shape, inputs, multipliers=1):
    hs.hs_circuit_record.restype = ctypes.c_void_p
    idx = np.array(inputs[0], dtype=np.uint32)
    vals = np.array(inputs[1], dtype=np.uint32)
    h = hs.hs_circuit_record(O.vp(shape), O.vp(idx), O.vp(vals), idx.size, multipliers)
    assert h, "recorder failed"
    h = ctypes.c_void_p(h)
    info = np.zeros(len(INFO), dtype=np.uint32)
    hs.hs_circuit_info(h, O.vp(info))

    def get(what, n):
        out = np.zeros(n, dtype=np.uint32)
        hs.hs_circuit_get(h, what, O.vp(out))
        return out
    return h, dict(zip(INFO, (int(x) for x in info))), get


def record_last(hs, shape):
    hs.hs_circuit_record_last.restype = ctypes.c_void_p
    h = hs.hs_circuit_record_last(O.vp(shape))
    assert h, "recorder failed"
    h = ctypes.c_void_p(h)
    info = np.zeros(len(INFO), dtype=np.uint32)
    hs.hs_circuit_info(h, O.vp(info))

    def get(what, n):
        out = np.zeros(n, dtype=np.uint32)
        hs.hs_circuit_get(h, what, O.vp(out))
        return out
    return h, dict(zip(INFO, (int(x) for x in info))), get


def evaluate(hs, h, info, blobs, shape, inputs, use_hints=0, full=1):
    hs.hs_verify_batch.restype = ctypes.c_void_p
    words, off = pack(blobs)
    idx = np.array(inputs[0], dtype=np.uint32)
    vals = np.array(inputs[1], dtype=np.uint32)
    n = len(blobs)
    dt = (Detail * n)()
    ws = np.zeros(4096, dtype=np.uint8)
    base = hs.hs_verify_batch(O.vp(words), O.vp(off), n, O.vp(shape), O.vp(idx), O.vp(vals), idx.size, full, dt, O.vp(ws))
    variables = np.zeros((n, info["n_vars"], 4), dtype=np.uint32)
    fh = np.zeros((n, info["n_flow"], 32), dtype=np.uint32)
    fs = np.zeros((n, info["n_flow"]), dtype=np.uint8)
    bad = np.zeros(n, dtype=np.int64)
    hs.hs_circuit_eval(h, O.vp(ws), n, O.vp(variables), O.vp(fh), O.vp(fs), None, O.vp(bad), use_hints)
    hs.hs_free(ctypes.c_void_p(base))
    return dt, variables, fh, fs, bad


@pytest.mark.parametrize("name,mult", [("small_proof.bin", 1), ("level13-1.bin", 1), ("level7-1.bin", 1), ("level9-1.bin", 1),
                                       ("small_proof.bin", 2)])
def test_recorded_circuit_matches_oracle(hostsim, orc, name, mult, monkeypatch):
    monkeypatch.setenv("STWO_B200_RECORDED_ORDER", "1")      # also build the second order of the tape (opt-in), checked below
    cs, out = oracle_circuit(name, mult)
    buf, n = O.load_proof(name)
    shape = shape_of(buf)
    h, info, get = record(hostsim, shape, O.inputs_for(name), mult)
    compare_wiring(cs, info, get)
    levels = get(10, info["n_levels"] + 1)
    assert levels[0] == 0 and levels[-1] == info["n_ins"] and np.all(np.diff(levels.astype(np.int64)) > 0)
    dt, variables, fh, fs, bad = evaluate(hostsim, h, info, [(buf, n)], shape, O.inputs_for(name))
    assert dt[0].verdict == 0 and bad[0] == -1
    want = np.array(cs.variables, dtype=np.uint32)
    diff = np.nonzero((variables[0] != want).any(axis=1))[0]
    assert diff.size == 0, "variable %d differs: %s != %s" % (diff[0], variables[0][diff[0]], want[diff[0]])
    # the same with every circuit permutation taken from the native verifier's record (tape::Perm::hint): identical
    # variables and flow prove the slot map -- transcript order, path order, the circuit's columns-before-node order
    assert hostsim.hs_circuit_hint_count(h) == info["n_flow"]          # every permutation of the circuit is one the native pass executes
    dt2, variables2, fh2, fs2, bad2 = evaluate(hostsim, h, info, [(buf, n)], shape, O.inputs_for(name), use_hints=1)
    assert bad2[0] == -1 and np.array_equal(variables2, variables) and np.array_equal(fh2, fh) and np.array_equal(fs2, fs)
    # ... and with the record produced by the cooperative tree rebuilds (every node hashed once, its states handed to the queries
    # whose path runs through it) instead of the per-query path stages: the product's default
    dt3, variables3, fh3, fs3, bad3 = evaluate(hostsim, h, info, [(buf, n)], shape, O.inputs_for(name), use_hints=1, full=3)
    assert dt3[0].n_perms_paths == out.n_perms_paths
    assert bad3[0] == -1 and np.array_equal(variables3, variables) and np.array_equal(fh3, fh) and np.array_equal(fs3, fs)
    # ... and in the SECOND ORDER of the tape (every recorded permutation split into outputs-from-the-record + flow entry: the
    # transcript and path chains are no dependency chains any more): fewer levels, the very same variables and flow
    o2 = np.zeros(3, dtype=np.uint32)
    hostsim.hs_circuit_recorded_order(h, O.vp(o2))
    assert 0 < o2[0] < info["n_levels"] and o2[2] == info["n_ins"] + info["n_flow"]
    dt4, variables4, fh4, fs4, bad4 = evaluate(hostsim, h, info, [(buf, n)], shape, O.inputs_for(name), use_hints=2, full=3)
    assert bad4[0] == -1 and np.array_equal(variables4, variables) and np.array_equal(fh4, fh) and np.array_equal(fs4, fs)
    wire, addr, wh, wsw = cs.flow_arrays()
    assert np.array_equal(fh[0], wh) and np.array_equal(fs[0], wsw)
    hostsim.hs_circuit_free(h)


def test_tampered_proof_fails_check_arithmetics(hostsim, orc):
    """a proof the native verifier rejects leaves an unsatisfied row in the circuit (the reference would panic on it)"""
    name = "small_proof.bin"
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    bad_buf = buf.copy()
    bad_buf[offs["sampled0"] + 1] ^= 4
    shape = shape_of(buf)
    h, info, get = record(hostsim, shape, O.inputs_for(name))
    dt, variables, fh, fs, bad = evaluate(hostsim, h, info, [(buf, n), (bad_buf, n)], shape, O.inputs_for(name))
    assert (dt[0].verdict, bad[0]) == (0, -1)
    assert dt[1].verdict == 1 and bad[1] >= 0
    hostsim.hs_circuit_free(h)


@pytest.mark.parametrize("name", ["level13-1.bin", "level10-1.bin"])
def test_last_layer_circuit_matches_oracle(hostsim, orc, name):
    """examples/last-layer (Plonk-without-Poseidon system, emulated Poseidon2): wiring and variables against the oracle"""
    cs, out = oracle_last_circuit(name)
    buf, n = O.load_proof(name)
    shape = shape_of(buf)
    h, info, get = record_last(hostsim, shape)
    compare_wiring_without(cs, info, get)
    dt, variables, fh, fs, bad = evaluate(hostsim, h, info, [(buf, n)], shape, O.inputs_for(name))
    assert dt[0].verdict == 0 and bad[0] == -1
    want = np.array(cs.variables, dtype=np.uint32)
    diff = np.nonzero((variables[0] != want).any(axis=1))[0]
    assert diff.size == 0, "variable %d differs: %s != %s" % (diff[0], variables[0][diff[0]], want[diff[0]])
    hostsim.hs_circuit_free(h)
