"""GPU tier: EVERY step of the reference's recursion chain (examples/single-proof/src/main.rs:91-101,
examples/multi-proofs/src/main.rs:173-296) traced on the device and compared with the committed oracle digests of
tests/golden/trace_digests.json -- including the shapes the other GPU tests leave out because the oracle's Python DSL is too
slow to re-run beside them: recursive_proof_16_15.bin verified five times inside one constraint system (2^19 rows, the
`multipliers` loop of multi-proofs main.rs:69-139), level1-5 / level4-5 (80 queries, 2^19 rows), level3-1 x 5, level6-1,
level8-1, level12-1.  Digest-only, so nothing under oracle/ has to run for 2^19-row circuits on the GPU box; the digests
themselves are recomputed from the oracle by tests/test_oracle_dsl.py on the CPU tier."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(O.ROOT, "tests", "golden", "trace_digests.json")))


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u4").tobytes()).hexdigest()


def _inputs(pkg, name):
    return pkg.INPUTS_SINGLE if name.startswith("small") else pkg.INPUTS_RECURSIVE


@pytest.mark.parametrize("g", GOLD["chain"], ids=["%s-x%d" % (g["src"][:-4], g["multipliers"]) for g in GOLD["chain"]])
def test_chain_step_matches_golden(pkg, gpu, g):
    name, mult = g["src"], g["multipliers"]
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    n = 34 if g["log_rows"] <= 17 else 3                # two lane groups with a ragged tail / one partial group for 2^19 rows
    vb = pkg.VerifyBatch([blob] * n, inputs=_inputs(pkg, name))
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, inputs=_inputs(pkg, name), multipliers=mult)
    i = circ.info
    assert (i.n_rows, i.n_rows_unpadded, i.n_flow, i.n_vars) == (1 << g["log_rows"], g["rows"], g["flow"], g["vars"])
    # the circuit hashes exactly what the native path stage hashes (per verification inside the constraint system)
    assert vb.fetch(n - 1, "detail").n_perms_paths * mult == g["flow"]
    assert _digest(circ.column("flow_wire").reshape(-1, 4)) == g["flow_wire_sha256"]
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    for p in (0, n - 1):
        assert _digest(pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])) == g["trace_sha256"], p
        assert _digest(circ.fetch(p, "flow_hash")) == g["flow_hash_sha256"], p


@pytest.mark.parametrize("g", GOLD["last_layer"], ids=[g["src"][:-4] for g in GOLD["last_layer"]])
def test_last_layer_matches_golden(pkg, gpu, g):
    """examples/last-layer on the three Poseidon31 shapes with a committed digest (level10-1: 2^18 rows)"""
    blob = open(os.path.join(O.PROOFS_DIR, g["src"]), "rb").read()
    n = 5
    vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_RECURSIVE)
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, last_layer=True)
    i = circ.info
    assert (i.n_rows, i.n_rows_unpadded, i.n_vars, i.num_input) == (1 << g["log_rows"], g["rows"], g["vars"], g["public_inputs"])
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all()
    for p in (0, n - 1):
        assert _digest(pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])) == g["trace_sha256"], p
