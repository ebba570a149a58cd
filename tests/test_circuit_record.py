"""CPU tier: the product's circuit recorder through the C ABI (host logic of libstwo_b200.so; no device needed) against the
oracle's circuit DSL -- every wiring column, row by row, for the shapes of the reference's fixtures."""
import numpy as np
import pytest

import oracle_py as O
from circuit_common import compare_wiring, oracle_circuit

CASES = [("small_proof.bin", 1), ("level13-1.bin", 1), ("level8-1.bin", 1), ("level6-1.bin", 1), ("small_proof.bin", 3)]


@pytest.mark.parametrize("name,mult", CASES)
def test_recorded_wiring_matches_oracle(pkg, orc, name, mult):
    cs, _ = oracle_circuit(name, mult)
    blob = open(O.PROOFS_DIR + "/" + name, "rb").read()
    circ = pkg.VerifierCircuit(pkg.proof_shape(blob), inputs=pkg.INPUTS_SINGLE if name.startswith("small") else pkg.INPUTS_RECURSIVE,
                               multipliers=mult)
    info = {k: getattr(circ.info, k) for k, _ in circ.info._fields_}
    names = {v: k for k, v in pkg._lib.COLUMNS.items()}
    compare_wiring(cs, info, lambda what, n: circ.column(names[what]))
    lv = circ.column("level_start").astype(np.int64)
    assert lv[0] == 0 and lv[-1] == info["n_ins"] and np.all(np.diff(lv) > 0)
    assert info["words_per_instance"] * mult == info["n_input_words"]


def test_record_rejects_bad_shapes(pkg):
    s = pkg.ProofShape()
    with pytest.raises(pkg.StwoB200Error):
        pkg.VerifierCircuit(s)
    blob = open(O.PROOFS_DIR + "/small_proof.bin", "rb").read()
    s = pkg.proof_shape(blob)
    s.n_queries = 0
    with pytest.raises(pkg.StwoB200Error):
        pkg.VerifierCircuit(s)


def test_trace_needs_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    blob = open(O.PROOFS_DIR + "/small_proof.bin", "rb").read()
    circ = pkg.VerifierCircuit(pkg.proof_shape(blob), inputs=pkg.INPUTS_SINGLE)
    rc = pkg._lib.load().stwo_b200_circuit_trace_batch_dev(circ._h, 1, 1, 1, 1, 1, 1 << 30, 0, None, None, None, None, None)
    assert rc == pkg._lib.E_NO_DEVICE


@pytest.mark.parametrize("name", ["level13-1.bin", "level12-1.bin"])
def test_recorded_last_layer_wiring_matches_oracle(pkg, orc, name):
    """examples/last-layer: Plonk-without-Poseidon wiring (wires + four selectors) through the C ABI"""
    from circuit_common import compare_wiring_without, oracle_last_circuit
    cs, _ = oracle_last_circuit(name)
    blob = open(O.PROOFS_DIR + "/" + name, "rb").read()
    circ = pkg.VerifierCircuit(pkg.proof_shape(blob), last_layer=True)
    info = {k: getattr(circ.info, k) for k, _ in circ.info._fields_}
    assert (info["kind"], info["n_preprocessed_columns"]) == (1, 8)
    names = {v: k for k, v in pkg._lib.COLUMNS.items()}
    compare_wiring_without(cs, info, lambda what, n: circ.column(names[what]))
    if name == "level13-1.bin":
        assert info["n_rows"] == 1 << 17            # header of examples/last-layer/data/bitcoin_proof.bin
