"""CPU tier: the oracle's full verifier pinned by the reference's own fixtures (embedded known answers: PoW low
bits, Merkle roots, logup sum, OODS equality, FRI last-layer equality) and by SURVEY App. F golden values."""
import json
import os

import numpy as np
import pytest

import oracle_py as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
POSEIDON_FIXTURES = sorted(f for f in os.listdir(O.PROOFS_DIR) if f.endswith(".bin") and f != "level14-1.bin")


def test_manifest_lists_every_fixture():
    man = json.load(open(os.path.join(O.PROOFS_DIR, "MANIFEST.json")))
    assert sorted(man) == sorted(POSEIDON_FIXTURES + ["level14-1.bin"])
    assert len(POSEIDON_FIXTURES) == 15


@pytest.mark.parametrize("name", POSEIDON_FIXTURES)
def test_every_poseidon31_fixture_is_accepted(orc, name):
    buf, n = O.load_proof(name)
    out = O.verify_proof(buf, n, O.inputs_for(name))
    assert (out.verdict, O.STAGES[out.stage]) == (0, "ok")
    pow_bits = int(np.frombuffer(buf[48:52].tobytes(), dtype=np.uint32)[0])
    assert out.digest_after_nonce[0] & ((1 << pow_bits) - 1) == 0
    assert list(out.oods_computed) == list(out.oods_expected)


def test_hybrid_hash_fixture_does_not_parse(orc):
    buf, n = O.load_proof("level14-1.bin")      # SHA-256 hybrid hasher: 32 raw bytes per hash, not 8 M31 words
    out = O.verify_proof(buf, n, O.INPUTS_RECURSIVE)
    assert (out.verdict, O.STAGES[out.stage]) == (1, "parse")


def test_survey_golden_values(orc):
    g = json.load(open(os.path.join(GOLD, "survey_appF.json")))
    s = g["small_proof.bin"]
    buf, n = O.load_proof("small_proof.bin")
    o = O.verify_proof(buf, n, O.INPUTS_SMALL)
    for k in ("z", "alpha", "random_coeff", "oods_t", "oods_x", "oods_y", "after_coeff"):
        assert list(getattr(o, k)) == s[k], k
    assert list(o.fri_alphas[0]) == s["fri_alpha_0"] and list(o.fri_alphas[7]) == s["fri_alpha_7"]
    assert list(o.digest_after_nonce) == s["digest_after_nonce"]
    assert list(o.log_sizes)[: o.n_logs] == [15, 13, 9]
    assert list(o.query_pos[0])[:16] == s["query_positions"]
    assert list(o.domain_points[0][0]) == s["domain_point_q0"]
    for gi, L in enumerate(("15", "13", "9")):
        assert list(o.fri_answers[gi][0]) == s["fri_answer_q0"][L]
        assert list(o.circle_folds[gi][0]) == s["circle_fold_q0"][L]
    assert o.n_transcript_perms == s["transcript_perms"] and o.n_perms_paths == s["total_path_perms"]
    r = g["recursive_proof_16_15.bin"]
    buf, n = O.load_proof("recursive_proof_16_15.bin")
    o = O.verify_proof(buf, n, O.INPUTS_RECURSIVE)
    for k in ("z", "alpha", "random_coeff", "oods_t", "after_coeff"):
        assert list(getattr(o, k)) == r[k], k
    assert list(o.fri_alphas[0]) == r["fri_alpha_0"] and o.digest_after_nonce[0] == r["digest_word0"]
    assert list(o.query_pos[0])[:4] == r["query_positions_first4"]
    assert o.n_transcript_perms == r["transcript_perms"] and o.n_perms_paths == r["total_path_perms"]


TAMPER = [("commitment0", "pow"), ("sampled0", "pow"), ("pow_nonce", "pow"), ("last_coeffs", "pow"), ("queried0", "merkle"),
          ("hash_witness0", "merkle"), ("queried3", "merkle"), ("fri_first_witness", "fri_first"),
          ("fri_first_hash_witness", "fri_first"), ("fri_inner0_witness", "fri_inner"), ("fri_inner0_hash_witness", "fri_inner"),
          ("fri_inner_last_witness", "fri_inner")]


@pytest.mark.parametrize("name", ["small_proof.bin", "level8-1.bin"])
@pytest.mark.parametrize("region,stage", TAMPER)
def test_single_bit_flips_are_rejected_at_the_expected_stage(orc, name, region, stage):
    buf, n = O.load_proof(name)
    off = O.proof_offsets(buf, n)[region]
    bad = buf.copy()
    bad[off] ^= 1
    out = O.verify_proof(bad, n, O.inputs_for(name))
    assert out.verdict == 1 and O.STAGES[out.stage] == stage


def test_wrong_public_inputs_fail_logup(orc):
    buf, n = O.load_proof("small_proof.bin")
    out = O.verify_proof(buf, n, O.INPUTS_RECURSIVE)
    assert (out.verdict, O.STAGES[out.stage]) == (1, "logup")


def test_truncated_and_garbage_blobs(orc):
    buf, n = O.load_proof("small_proof.bin")
    for cut in (0, 3, 100, n - 4):
        assert O.STAGES[O.verify_proof(buf, cut, O.INPUTS_SMALL).stage] == "parse"
    junk = np.full(4096, 0xFF, dtype=np.uint8)
    assert O.STAGES[O.verify_proof(junk, 4096, O.INPUTS_SMALL).stage] == "parse"
