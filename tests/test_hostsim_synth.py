"""CPU tier: the synthetic FRI + Merkle instances of BASELINE configs[4] part i (recursive-stwo_b200/csrc/synth.cuh).  The generator
(seeded low-degree columns -> mixed-degree commit -> channel -> folds + commits -> last-layer polynomial -> queries -> stwo-layout
decommitments) and the product's verifier stages run here compiled for the host; the oracle's FRI-only verifier (oracle/orc_verify.c,
orc_fri_verify_synth = the reference's FRI query phase, fri_stage) is the checker: every instance must be accepted by both, every
fold / path root must agree, and a flipped word must be rejected by both at the same stage."""
import ctypes

import numpy as np
import pytest

import oracle_py as O
from verify_common import Detail

# (log_size_plonk, log_size_poseidon, pow_bits, log_blowup, log_last, n_queries, n_inner): max_first = max(lsp + 1, lspos + 2) + blowup
SHAPES = {"three-sizes": (3, 4, 3, 2, 1, 5, 4), "two-sizes": (5, 4, 0, 1, 2, 7, 3), "deep": (2, 6, 2, 3, 0, 9, 7)}


def generate(hs, shape, seed):
    hs.hs_synth_blob_words.restype = ctypes.c_uint32
    sh = np.array(shape, dtype=np.uint32)
    n = hs.hs_synth_blob_words(O.vp(sh))
    blob = np.zeros(n, dtype=np.uint32)
    rc = hs.hs_synth_generate(O.vp(sh), ctypes.c_uint64(seed), O.vp(blob))
    assert rc == 0, "generator self-check failed: %d" % rc
    return blob


def product_verify(hs, blobs, shape, coop):
    hs.hs_synth_verify.restype = ctypes.c_void_p
    sh = np.array(shape, dtype=np.uint32)
    words = np.concatenate(blobs)
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([b.size for b in blobs])
    dt = (Detail * len(blobs))()
    base = hs.hs_synth_verify(O.vp(words), O.vp(off), len(blobs), O.vp(sh), coop, dt, None)
    hs.hs_free(ctypes.c_void_p(base))
    return dt


def oracle_verify(orc, blob):
    out = O.VerifyOut()
    orc.orc_fri_verify_synth(O.vp(blob), ctypes.c_size_t(blob.size), ctypes.byref(out))
    return out


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("coop", [0, 1], ids=["thread-per-tree", "cooperative"])
def test_instances_are_accepted_by_product_and_oracle(hostsim, orc, name, coop):
    shape = SHAPES[name]
    blobs = [generate(hostsim, shape, seed) for seed in range(6)]
    assert len({b.tobytes() for b in blobs}) == 6                         # distinct instances
    dts = product_verify(hostsim, blobs, shape, coop)
    seen_positions = set()
    for b, dt in zip(blobs, dts):
        o = oracle_verify(orc, b)
        assert (o.verdict, o.stage) == (0, 0)
        assert (dt.verdict, dt.stage) == (0, 0), bin(dt.fail_mask)
        nq, n_inner = shape[5], shape[6]
        assert list(dt.fs.raw_queries)[:nq] == list(o.raw_queries)[:nq]
        for i in range(n_inner + 1):
            assert list(dt.fs.fri_alphas[i]) == list(o.fri_alphas[i])
        assert dt.n_perms_hints == o.n_perms_hints and dt.n_perms_paths == o.n_perms_paths
        seen_positions.add(tuple(o.raw_queries)[:nq])
    assert len(seen_positions) == 6                                       # every instance opens its own positions


@pytest.mark.parametrize("name", ["three-sizes", "deep"])
def test_tampered_instances_are_rejected_alike(hostsim, orc, name):
    shape = SHAPES[name]
    blob = generate(hostsim, shape, 77)
    hdr = blob[:256]
    sections = {"fl_commitment": int(hdr[80]), "last_coeffs": int(hdr[81]), "answers": int(hdr[82]), "fl_fri_witness": int(hdr[83]),
                "fl_hash_witness": int(hdr[84]), "in_commitment0": int(hdr[96]), "in_hash_witness0": int(hdr[160]),
                "in_hash_witness_last": int(hdr[160 + shape[6] - 1]), "in_fri_witness1": int(hdr[128 + 1])}
    blobs = [blob]
    for sec, off in sections.items():
        if sec == "fl_fri_witness" and hdr[8] == 0 or sec == "in_fri_witness1" and hdr[10 + 1] == 0:
            continue
        b = blob.copy()
        b[off] ^= 1
        blobs.append(b)
    b = blob.copy(); b[8] += 1; blobs.append(b)                          # one witness value too many claimed
    b = blob.copy(); b[42] = max(int(b[42]), 1) - 1; blobs.append(b)      # one hash witness too few
    stages = set()
    for coop in (0, 1):
        dts = product_verify(hostsim, blobs, shape, coop)
        for b, dt in zip(blobs, dts):
            o = oracle_verify(orc, b)
            assert (dt.verdict, dt.stage) == (o.verdict, o.stage), (dt.verdict, dt.stage, o.verdict, o.stage)
            stages.add(O.STAGES[o.stage])
    assert {"ok", "fri_first", "fri_inner"} <= stages and ("fri_last" in stages or "pow" in stages)


# ---- the folding-stage circuit over an instance (the tape of configs[4] part i) ----------------------------------------------------------
INFO = ("n_rows", "n_rows_unpadded", "n_vars", "n_flow", "n_flow_padded", "n_input_words", "n_ins", "n_levels", "num_input", "words_per_instance")


@pytest.mark.parametrize("name", ["three-sizes", "two-sizes", "deep"])
def test_folding_circuit_matches_oracle(hostsim, orc, name):
    """FoldingResults::compute alone, everything before it as witnesses: the product's recorder + tape evaluator against the oracle DSL's
    `folding` fed by the oracle's FRI-only verifier -- same wiring, same variables, same Poseidon flow; and the same again with every
    permutation taken from the record the product's tree rebuilds leave behind."""
    from circuit_common import D, compare_wiring
    shape = SHAPES[name]
    blobs = [generate(hostsim, shape, seed) for seed in (3, 4)]
    cs, out = D.folding_circuit(blobs[0], O.VerifyOut)
    hostsim.hs_circuit_record_folding.restype = ctypes.c_void_p
    sh = np.array(shape, dtype=np.uint32)
    h = ctypes.c_void_p(hostsim.hs_circuit_record_folding(O.vp(sh)))
    assert h, "recorder failed"
    info = np.zeros(len(INFO), dtype=np.uint32)
    hostsim.hs_circuit_info(h, O.vp(info))
    info = dict(zip(INFO, (int(x) for x in info)))

    def get(what, n):
        o = np.zeros(n, dtype=np.uint32)
        hostsim.hs_circuit_get(h, what, O.vp(o))
        return o
    compare_wiring(cs, info, get)
    assert hostsim.hs_circuit_hint_count(h) == info["n_flow"] == len(cs.flow)
    hostsim.hs_synth_verify.restype = ctypes.c_void_p
    words = np.concatenate(blobs)
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([b.size for b in blobs])
    results = []
    for coop, use_hints in ((1, 0), (1, 1), (0, 1)):
        dt = (Detail * len(blobs))()
        ws = np.zeros(4096, dtype=np.uint8)
        base = hostsim.hs_synth_verify(O.vp(words), O.vp(off), len(blobs), O.vp(sh), coop, dt, O.vp(ws))
        assert [(d.verdict, d.stage) for d in dt] == [(0, 0)] * len(blobs)
        variables = np.zeros((len(blobs), info["n_vars"], 4), dtype=np.uint32)
        fh = np.zeros((len(blobs), info["n_flow"], 32), dtype=np.uint32)
        fsw = np.zeros((len(blobs), info["n_flow"]), dtype=np.uint8)
        bad = np.zeros(len(blobs), dtype=np.int64)
        hostsim.hs_circuit_eval(h, O.vp(ws), len(blobs), O.vp(variables), O.vp(fh), O.vp(fsw), None, O.vp(bad), use_hints)
        hostsim.hs_free(ctypes.c_void_p(base))
        assert list(bad) == [-1, -1]
        results.append((variables, fh, fsw))
    want = np.array(cs.variables, dtype=np.uint32)
    diff = np.nonzero((results[0][0][0] != want).any(axis=1))[0]
    assert diff.size == 0, "variable %d differs: %s != %s" % (diff[0], results[0][0][0][diff[0]], want[diff[0]])
    wire, addr, wh, wsw = cs.flow_arrays()
    assert np.array_equal(results[0][1][0], wh) and np.array_equal(results[0][2][0], wsw)
    for r in results[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(r, results[0]))
    # the second instance satisfies the same circuit with its own values
    cs2, _ = D.folding_circuit(blobs[1], O.VerifyOut)
    assert np.array_equal(results[0][0][1], np.array(cs2.variables, dtype=np.uint32))
    hostsim.hs_circuit_free(h)


def test_folding_circuit_of_a_tampered_instance_is_unsatisfied(hostsim, orc):
    shape = SHAPES["three-sizes"]
    good = generate(hostsim, shape, 5)
    bad_blob = good.copy()
    bad_blob[int(good[82])] ^= 1                                           # first answer word
    sh = np.array(shape, dtype=np.uint32)
    hostsim.hs_circuit_record_folding.restype = ctypes.c_void_p
    h = ctypes.c_void_p(hostsim.hs_circuit_record_folding(O.vp(sh)))
    info = np.zeros(len(INFO), dtype=np.uint32)
    hostsim.hs_circuit_info(h, O.vp(info))
    info = dict(zip(INFO, (int(x) for x in info)))
    hostsim.hs_synth_verify.restype = ctypes.c_void_p
    words = np.concatenate([good, bad_blob])
    off = np.array([0, good.size, 2 * good.size], dtype=np.uint64)
    dt = (Detail * 2)()
    ws = np.zeros(4096, dtype=np.uint8)
    base = hostsim.hs_synth_verify(O.vp(words), O.vp(off), 2, O.vp(sh), 1, dt, O.vp(ws))
    variables = np.zeros((2, info["n_vars"], 4), dtype=np.uint32)
    fh = np.zeros((2, info["n_flow"], 32), dtype=np.uint32)
    fsw = np.zeros((2, info["n_flow"]), dtype=np.uint8)
    bad = np.zeros(2, dtype=np.int64)
    hostsim.hs_circuit_eval(h, O.vp(ws), 2, O.vp(variables), O.vp(fh), O.vp(fsw), None, O.vp(bad), 1)
    hostsim.hs_free(ctypes.c_void_p(base))
    assert (dt[0].verdict, bad[0]) == (0, -1)
    assert dt[1].verdict == 1 and bad[1] >= 0                              # the reference would panic inside equalverify
    hostsim.hs_circuit_free(h)
