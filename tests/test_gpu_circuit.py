"""GPU tier: trace generation of the recursive verifier circuit through the C ABI (gather -> K6 tape evaluation -> K7 checks
and export), bit-exact against the oracle's circuit DSL: variables[], Poseidon flow, the 22 trace columns (golden digests of
tests/golden/trace_digests.json), check_arithmetics / check_poseidon_invocations verdicts on accepted and tampered proofs."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle_py as O
from circuit_common import D, oracle_circuit

pytestmark = pytest.mark.gpu
GOLD = {g["src"]: g for g in json.load(open(os.path.join(O.ROOT, "tests", "golden", "trace_digests.json")))["chain"]}


def _inputs(pkg, name):
    return pkg.INPUTS_SINGLE if name.startswith("small") else pkg.INPUTS_RECURSIVE


@pytest.mark.parametrize("name", ["small_proof.bin", "level13-1.bin", "level7-1.bin", "level9-1.bin", "level11-1.bin"])
def test_trace_matches_oracle_and_golden(pkg, gpu, orc, name):
    cs, out = oracle_circuit(name, 1)
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    n = 40                                           # more than one lane group, ragged tail
    vb = pkg.VerifyBatch([blob] * n, inputs=_inputs(pkg, name))
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, inputs=_inputs(pkg, name))
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    want = np.array(cs.variables, dtype=np.uint32)
    wire, addr, wh, wsw = cs.flow_arrays()
    for p in (0, 31, 32, n - 1):
        assert np.array_equal(circ.fetch(p, "variables"), want)
        assert np.array_equal(circ.fetch(p, "flow_hash"), wh) and np.array_equal(circ.fetch(p, "flow_swap"), wsw)
    tr = cs.trace_columns()
    for p in (0, 33, n - 1):
        got = pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])
        bad = np.argwhere(got != tr)
        assert bad.size == 0, "column %s row %d" % (pkg.circuit.COLUMN_NAMES[bad[0][0]], bad[0][1])
        assert D.trace_digest(got) == GOLD[name]["trace_sha256"]


@pytest.mark.parametrize("pair", [("level10-1.bin", "level11-1.bin"), ("level2-1.bin", "level5-1.bin")])
def test_different_proofs_in_neighbouring_lanes(pkg, gpu, orc, pair):
    """one batch, one shape, two different proofs alternating lane by lane (the second pair also differs in byte length):
    every proof's trace equals the golden digest of its own fixture"""
    blobs = [open(os.path.join(O.PROOFS_DIR, f), "rb").read() for f in pair]
    n = 70
    vb = pkg.VerifyBatch([blobs[k % 2] for k in range(n)], inputs=pkg.INPUTS_RECURSIVE)
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_RECURSIVE)
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    for p in (0, 1, 31, 32, 63, 64, n - 1):
        got = pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])
        assert D.trace_digest(got) == GOLD[pair[p % 2]]["trace_sha256"], p


def test_native_hints_and_recomputation_agree(pkg, gpu, orc):
    """K6 with the circuit's permutations taken from the native verifier's record (the default after run(full=True)) and with
    every permutation executed again: identical variables, flow and trace; without full mode the hints are not used"""
    blob = open(os.path.join(O.PROOFS_DIR, "small_proof.bin"), "rb").read()
    vb = pkg.VerifyBatch([blob] * 37, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    got = {}
    for mode in (True, False):
        r = circ.trace(vb, check=True, export=True, native_hints=mode)
        assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
        got[mode] = (circ.fetch(36, "variables").copy(), circ.fetch(36, "flow_hash").copy(), r["values"].cpu().numpy().copy())
    for a, b in zip(got[True], got[False]):
        assert np.array_equal(a, b)
    vb.run(full=False)                                  # no per-query paths: nothing to reuse
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and np.array_equal(r["values"].cpu().numpy(), got[True][2])


@pytest.mark.parametrize("name,n", [("small_proof.bin", 300), ("level2-1.bin", 9)])
def test_repeated_runs_are_bit_identical(pkg, gpu, orc, name, n):
    """the cooperative stages, the stream pool and the grid-wide tape evaluation leave no room for a race: three runs of the same
    batch (sliced path for 300 proofs) give the same workspace-derived values, variables and trace, bit for bit"""
    import hashlib
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    vb = pkg.VerifyBatch([blob] * n, inputs=_inputs(pkg, name))
    circ = pkg.VerifierCircuit(vb.shape, inputs=_inputs(pkg, name))
    seen = set()
    for _ in range(3):
        v, s = vb.run(full=True)
        r = circ.trace(vb, check=True, export=True, preprocessed=False)
        assert not v.cpu().numpy().any() and (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
        h = hashlib.sha256()
        h.update(r["values"].cpu().numpy().tobytes())
        h.update(circ.fetch(n - 1, "variables").tobytes())
        for what in ("answers", "line_folds", "path_roots", "pair_hints"):
            h.update(vb.fetch(n // 2, what).tobytes())
        seen.add(h.hexdigest())
    assert len(seen) == 1


def test_multipliers(pkg, gpu, orc):
    """examples/multi-proofs: the same proof verified twice inside one constraint system"""
    name = "small_proof.bin"
    cs, _ = oracle_circuit(name, 2)
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    vb = pkg.VerifyBatch([blob] * 3, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE, multipliers=2)
    r = circ.trace(vb)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    got = pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][2])
    assert np.array_equal(got, cs.trace_columns())


def test_tampered_proofs_fail_check_arithmetics(pkg, gpu, orc):
    """a rejected proof leaves an unsatisfied row (the reference would panic inside the DSL); accepted neighbours are unaffected"""
    name = "small_proof.bin"
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    blobs, expect_ok = [], []
    for k, region in enumerate(["sampled0", None, "queried0", "fri_first_witness", None, "last_coeffs", "pow_nonce", None]):
        b = buf.copy()
        if region:
            b[offs[region] + 1] ^= 1 << (k % 7)
        blobs.append(bytes(b[:n]))
        expect_ok.append(region is None)
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    verdict, _ = vb.run(full=True)
    verdict = verdict.cpu().numpy()
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    r = circ.trace(vb, export=False, preprocessed=False)
    bad_row = r["bad_row"].cpu().numpy()
    for k, ok in enumerate(expect_ok):
        assert (verdict[k] == 0) == ok
        assert (bad_row[k] == -1) == ok, (k, bad_row[k])
    assert (r["bad_flow"].cpu().numpy() == -1).all()     # the Poseidon flow is self-consistent whatever the proof says


def test_cs_finalize_host_entry(pkg, gpu, orc):
    """stwo_b200_cs_finalize: one constraint system whose values were produced on the host (here: by the oracle)"""
    cs, _ = oracle_circuit("small_proof.bin", 1)
    L = pkg._lib
    cols = {k: np.ascontiguousarray(getattr(cs, k), dtype=np.uint32) for k in ("a_wire", "b_wire", "c_wire", "poseidon_wire", "enforce_c_m31", "op")}
    wire, addr, wh, wsw = cs.flow_arrays()
    variables = np.ascontiguousarray(cs.variables, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    w = L.CsWiring(len(cs.variables), len(cs.a_wire), len(cs.flow), cs.num_input, *[vp(cols[k]) for k in cols], None, vp(wire), vp(addr))
    v = L.CsValues(1, 1, vp(variables), vp(wh), vp(wsw))
    trace = np.zeros((22, len(cs.a_wire)), dtype=np.uint32)
    bad_row, bad_flow = ctypes.c_int64(7), ctypes.c_int64(7)
    L.call("stwo_b200_cs_finalize", ctypes.byref(w), ctypes.byref(v), vp(trace), ctypes.byref(bad_row), ctypes.byref(bad_flow))
    assert (bad_row.value, bad_flow.value) == (-1, -1)
    assert np.array_equal(trace, cs.trace_columns())
    # a corrupted variable and a corrupted flow hash are located
    row = next(i for i in range(2000, len(cs.a_wire)) if cs.op[i] == 0 and cs.c_wire[i] > 3)
    variables[cs.c_wire[row], 0] ^= 1
    wh[17, 20] ^= 1
    L.call("stwo_b200_cs_finalize", ctypes.byref(w), ctypes.byref(v), vp(trace), ctypes.byref(bad_row), ctypes.byref(bad_flow))
    assert 0 <= bad_row.value <= row and bad_flow.value == 17


@pytest.mark.parametrize("name", ["level13-1.bin", "level12-1.bin"])
def test_last_layer_trace_matches_oracle(pkg, gpu, orc, name):
    """BASELINE configs[1] (examples/last-layer): the 20-column trace of the last-layer circuit, batch of replicas"""
    from circuit_common import oracle_last_circuit
    cs, out = oracle_last_circuit(name)
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    n = 35
    vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_RECURSIVE)
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, last_layer=True)
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    want = np.array(cs.variables, dtype=np.uint32)
    for p in (0, 32, n - 1):
        got = circ.fetch(p, "variables")
        diff = np.nonzero((got != want).any(axis=1))[0]
        assert diff.size == 0, "variable %d of proof %d" % (diff[0], p)
    tr = cs.trace_columns()
    gold = {g["src"]: g for g in json.load(open(os.path.join(O.ROOT, "tests", "golden", "trace_digests.json")))["last_layer"]}
    for p in (1, n - 1):
        got = pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])
        bad = np.argwhere(got != tr)
        assert bad.size == 0, "column %s row %d" % (pkg.circuit.COLUMN_NAMES_WITHOUT[bad[0][0]], bad[0][1])
        assert D.trace_digest(got) == gold[name]["trace_sha256"]
    # a tampered proof breaks a row of this circuit too
    buf, ln = O.load_proof(name)
    offs = O.proof_offsets(buf, ln)
    b = buf.copy()
    b[offs["sampled0"] + 2] ^= 8
    vb2 = pkg.VerifyBatch([blob, bytes(b[:ln])], inputs=pkg.INPUTS_RECURSIVE)
    v2, _ = vb2.run(full=True)
    r2 = circ.trace(vb2, export=False, preprocessed=False)
    assert v2.cpu().numpy().tolist() == [0, 1]
    br = r2["bad_row"].cpu().numpy()
    assert br[0] == -1 and br[1] >= 0


def test_recorded_check_equals_reexecution(pkg, gpu, orc):
    """check_poseidon_invocations against the record of executed permutations (entry input == recorded input, entry output == recorded
    output: the default after run(full=True)) and by executing every flow entry again (recheck=True): same verdict per proof, for
    accepted proofs and for rejected ones (whose record is incomplete, so both ways re-execute)"""
    name = "small_proof.bin"
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    blobs = []
    for k, region in enumerate([None, "queried0", None, "fri_first_witness", "sampled0", None] * 7):
        b = buf.copy()
        if region:
            b[offs[region] + 1] ^= 1 << (k % 5)
        blobs.append(bytes(b[:n]))
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    got = {}
    for recheck in (False, True):
        r = circ.trace(vb, check=True, export=True, recheck=recheck)
        got[recheck] = (r["bad_row"].cpu().numpy().copy(), r["bad_flow"].cpu().numpy().copy(), r["values"].cpu().numpy().copy())
    for a, b in zip(got[False], got[True]):
        assert np.array_equal(a, b)
    assert (got[False][1] == -1).all() and (got[False][0] == -1).sum() == 21


def test_recorded_check_is_live(pkg, gpu, orc):
    """the record-based check really compares: one flipped word in the INPUT record of a path permutation of one proof is reported as
    that proof's first bad flow entry (the evaluation never reads the input record, so every row still holds); the re-execution does
    not look at the record and stays clean"""
    import torch
    blob = open(os.path.join(O.PROOFS_DIR, "small_proof.bin"), "rb").read()
    n = 34
    vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    r = circ.trace(vb, check=True, export=False, preprocessed=False)
    assert (r["bad_flow"].cpu().numpy() == -1).all()
    # locate proof 33's input record inside the workspace through a sentinel: fetch, find the words, flip one on the device
    rec_in = vb.fetch(33, "perm_record_inputs")
    slot = 512 + 700                                                       # a path slot of the second commitment tree
    ws32 = vb.d_ws[: vb.d_ws.numel() // 4 * 4].view(torch.int32)
    needle = torch.from_numpy(rec_in[slot].view(np.int32).copy()).to(ws32.device)
    cand = (ws32[:-16] == needle[0]).nonzero().flatten()
    hits = [int(h) for h in cand.tolist() if bool((ws32[h: h + 16] == needle).all())]
    mine = None
    for h in hits:                                                         # replicas: every proof holds the same 16 words; take proof 33's
        ws32[h + 3] ^= 1
        if vb.fetch(33, "perm_record_inputs")[slot][3] != rec_in[slot][3]:
            mine = h
            break
        ws32[h + 3] ^= 1
    assert mine is not None
    r = circ.trace(vb, check=True, export=False, preprocessed=False)
    bad_flow, bad_row = r["bad_flow"].cpu().numpy(), r["bad_row"].cpu().numpy()
    assert bad_flow[33] >= 0 and (np.delete(bad_flow, 33) == -1).all() and (bad_row == -1).all()
    r = circ.trace(vb, check=True, export=False, preprocessed=False, recheck=True)
    assert (r["bad_flow"].cpu().numpy() == -1).all()


def test_second_tape_order_is_bit_identical(pkg, gpu, orc, monkeypatch):
    """the opt-in second order of the tape (STWO_B200_RECORDED_ORDER=1: every recorded permutation split into outputs-from-the-record +
    flow entry, 74 -> 55 levels) gives the very same variables, flow and trace; a lane group holding a rejected proof falls back to the
    first order on its own"""
    name = "small_proof.bin"
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    bad = buf.copy()
    bad[offs["queried0"] + 1] ^= 2
    blobs = [bytes(buf[:n])] * 40 + [bytes(bad[:n])] + [bytes(buf[:n])] * 9          # group 0 clean, group 1 holds the rejected proof
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    base = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    r0 = base.trace(vb, check=True, export=True)
    keep = (r0["values"].cpu().numpy().copy(), r0["bad_row"].cpu().numpy().copy(), r0["bad_flow"].cpu().numpy().copy(),
            base.fetch(3, "variables").copy(), base.fetch(3, "flow_hash").copy())
    monkeypatch.setenv("STWO_B200_RECORDED_ORDER", "1")
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)                     # recorded under the switch: both orders
    r1 = circ.trace(vb, check=True, export=True)
    got = (r1["values"].cpu().numpy(), r1["bad_row"].cpu().numpy(), r1["bad_flow"].cpu().numpy(), circ.fetch(3, "variables"), circ.fetch(3, "flow_hash"))
    for a, b in zip(keep, got):
        assert np.array_equal(a, b)
    assert keep[1][40] >= 0 and (np.delete(keep[1], 40) == -1).all()
