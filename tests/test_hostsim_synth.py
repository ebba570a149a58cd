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
