"""CPU tier: the product's verifier stage functions (verify.cuh and below), compiled for the host, against the oracle on
every fixture and on tampered proofs — the same code the CUDA kernels dispatch one thread per proof / tree / query."""
import ctypes
import os

import numpy as np
import pytest

import oracle_py as O
from verify_common import Detail, compare_detail, pack, shape_of

FIXTURES = sorted(f for f in os.listdir(O.PROOFS_DIR) if f.endswith(".bin") and f != "level14-1.bin")


def run_hostsim(hs, blobs, shape, inputs, full=1):
    hs.hs_verify_batch.restype = ctypes.c_void_p
    words, off = pack(blobs)
    idx = np.array(inputs[0], dtype=np.uint32)
    vals = np.array(inputs[1], dtype=np.uint32)
    dt = (Detail * len(blobs))()
    base = hs.hs_verify_batch(O.vp(words), O.vp(off), len(blobs), O.vp(shape), O.vp(idx), O.vp(vals), idx.size, full, dt, None)
    hs.hs_free(ctypes.c_void_p(base))
    return dt


def test_detail_struct_layout(hostsim):
    hostsim.hs_detail_size.restype = ctypes.c_size_t
    assert hostsim.hs_detail_size() == ctypes.sizeof(Detail)


@pytest.mark.parametrize("coop", [0, 2], ids=["thread-per-tree", "cooperative"])
@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_matches_oracle(hostsim, orc, name, coop):
    buf, n = O.load_proof(name)
    o = O.verify_proof(buf, n, O.inputs_for(name))
    dt = run_hostsim(hostsim, [(buf, n)], shape_of(buf), O.inputs_for(name), full=1 | coop)
    compare_detail(dt[0], o)
    assert dt[0].verdict == 0 and dt[0].n_perms_hints == o.n_perms_hints


REGIONS = ["commitment0", "sampled0", "pow_nonce", "last_coeffs", "queried0", "hash_witness0", "queried3", "fri_first_witness",
           "fri_first_hash_witness", "fri_inner0_witness", "fri_inner0_hash_witness", "fri_inner_last_witness"]


@pytest.mark.parametrize("coop", [0, 2], ids=["thread-per-tree", "cooperative"])
@pytest.mark.parametrize("name", ["small_proof.bin", "level13-1.bin"])
def test_tampered_batch_matches_oracle(hostsim, orc, name, coop):
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    blobs = [(buf, n)]
    for r in REGIONS:
        for bit in (0, 9):
            bad = buf.copy()
            bad[offs[r] + bit // 8] ^= 1 << (bit % 8)
            blobs.append((bad, n))
    blobs.append((buf, n - 4))             # truncated
    blobs.append((buf[:64].copy(), 64))    # header only
    dts = run_hostsim(hostsim, blobs, shape_of(buf), O.inputs_for(name), full=1 | coop)
    stages = set()
    for (b, ln), dt in zip(blobs, dts):
        o = O.verify_proof(b, ln, O.inputs_for(name))
        compare_detail(dt, o, full=False)
        stages.add(O.STAGES[o.stage])
    assert {"ok", "parse", "pow", "merkle", "fri_first", "fri_inner"} <= stages


def test_wrong_inputs_and_wrong_shape(hostsim, orc):
    buf, n = O.load_proof("small_proof.bin")
    dt = run_hostsim(hostsim, [(buf, n)], shape_of(buf), O.INPUTS_RECURSIVE)
    assert (dt[0].verdict, O.STAGES[dt[0].stage]) == (1, "logup")
    other, _ = O.load_proof("recursive_proof_16_15.bin")
    dt = run_hostsim(hostsim, [(buf, n)], shape_of(other), O.INPUTS_SMALL)       # batch shape != proof shape
    assert (dt[0].verdict, O.STAGES[dt[0].stage]) == (1, "parse")


# ---- the caller's PcsConfig, adversarial headers, the reference's panic case ------------------------------------------------
def _cfg(shape):
    return tuple(int(x) for x in shape[2:6])            # (pow_bits, log_blowup, log_last, n_queries)


def test_weak_config_header_is_rejected(hostsim, orc):
    """A proof is verified under the CALLER's PcsConfig (components/hints/src/fiat_shamir.rs:69-73), never under the one its
    header claims: small_proof.bin with pow_bits rewritten to 0 (its PoW check would then pass trivially), with fewer queries or
    with another blow-up is rejected at the parse stage by the product (batch shape = caller's) and by the oracle alike."""
    buf, n = O.load_proof("small_proof.bin")
    shape = shape_of(buf)
    w = buf.view(np.uint32)
    blobs = [(buf, n)]
    for word, val in ((10, 0), (10, 19), (13, 1), (13, 15), (11, 1), (11, 4), (12, 1)):     # pow_bits, n_queries, log_blowup, log_last
        bad = buf.copy()
        bad.view(np.uint32)[word] = val
        blobs.append((bad, n))
    dts = run_hostsim(hostsim, blobs, shape, O.INPUTS_SMALL)
    assert (dts[0].verdict, dts[0].stage) == (0, 0)
    for (b, ln), dt in list(zip(blobs, dts))[1:]:
        o = O.verify_proof(b, ln, O.INPUTS_SMALL, config=_cfg(shape))
        assert (o.verdict, O.STAGES[o.stage]) == (1, "parse")
        assert (dt.verdict, O.STAGES[dt.stage]) == (1, "parse")
    # the oracle run WITHOUT a caller config shows why: with pow_bits = 0 in its header the forged blob is accepted
    assert O.verify_proof(blobs[1][0], n, O.INPUTS_SMALL).verdict == 0


@pytest.mark.parametrize("word", [0, 1, 11], ids=["log_size_plonk", "log_size_poseidon", "log_blowup"])
@pytest.mark.parametrize("val", [0xFFFFFFFF, 0xFFFFFFFB, 29, 0x80000010, 0])
def test_garbage_header_words_do_not_wrap(hostsim, orc, word, val):
    """0xFFFFFFFF-style header fields: the u32 sums log_size + log_blowup must not wrap into a plausible shape (a wrapped
    log_size_plonk would send the OODS stage into a 2^32-step loop); rejected at parse by the product and the oracle"""
    buf, n = O.load_proof("small_proof.bin")
    shape = shape_of(buf)
    bad = buf.copy()
    bad.view(np.uint32)[word] = val
    dt = run_hostsim(hostsim, [(bad, n)], shape, O.INPUTS_SMALL)
    assert (dt[0].verdict, O.STAGES[dt[0].stage]) == (1, "parse")
    # ... and the shape the forged header claims is refused by the library before any workspace is sized for it
    import importlib
    import ctypes as C
    L = importlib.import_module("recursive-stwo_b200")._lib
    claimed = shape.copy()
    claimed[{0: 0, 1: 1, 11: 3}[word]] = val
    assert L.load().stwo_b200_verify_workspace_bytes(C.cast(O.vp(claimed), C.POINTER(L.ProofShape)), 4) == 0
    out = L.ProofShape()
    assert L.load().stwo_b200_proof_shape_of(O.vp(bad), n, C.byref(out)) == L.E_SHAPE
    o = O.verify_proof(bad, n, O.INPUTS_SMALL)
    assert (o.verdict, O.STAGES[o.stage]) == (1, "parse")


def test_fri_layer_count_is_tied_to_the_log_sizes(hostsim, orc):
    """stwo's InvalidNumFriLayers: every fixture satisfies max_first == max(log_size_plonk + 1, log_size_poseidon + 2) + blowup;
    a header whose log sizes were moved (so the claimed composition degree bound no longer matches the FRI layers) is rejected"""
    for name in FIXTURES:
        buf, n = O.load_proof(name)
        s = shape_of(buf)
        assert s[4] + s[3] + 1 + s[6] == max(s[0] + 1, s[1] + 2) + s[3], name
    buf, n = O.load_proof("recursive_proof_16_15.bin")
    shape = shape_of(buf)
    for word, val in ((0, 17), (1, 16), (1, 14), (0, 15)):
        bad = buf.copy()
        bad.view(np.uint32)[word] = val
        claimed = shape.copy()
        claimed[word] = val
        o = O.verify_proof(bad, n, O.INPUTS_RECURSIVE)
        dt = run_hostsim(hostsim, [(bad, n)], claimed, O.INPUTS_RECURSIVE)
        if max(int(claimed[0]) + 1, int(claimed[1]) + 2) == 17:      # (0, 15): the Poseidon component still sets the bound -> reaches the PoW check
            assert O.STAGES[o.stage] != "parse" and o.verdict == 1 and dt[0].stage == o.stage
        else:
            assert (o.verdict, O.STAGES[o.stage]) == (1, "parse") and (dt[0].verdict, O.STAGES[dt[0].stage]) == (1, "parse")


def test_duplicate_queries_are_unsupported(hostsim, orc):
    """VERDICT_UNSUPPORTED: the reference panics on duplicated queries at the largest domain
    (components/recursive/answer/src/lib.rs:190-195); forged query draws reach stage_after_transcript directly"""
    buf, n = O.load_proof("small_proof.bin")
    shape = shape_of(buf)
    hostsim.hs_after_transcript_verdict.restype = ctypes.c_uint32
    o = O.verify_proof(buf, n, O.INPUTS_SMALL)
    raw = np.array(list(o.raw_queries)[:16], dtype=np.uint32)
    assert hostsim.hs_after_transcript_verdict(O.vp(shape), O.vp(raw), 1) == 0                   # accept so far
    dup = raw.copy()
    dup[11] = dup[3] ^ (1 << 20)            # same position at log size 15 (max_first), different raw draw
    assert hostsim.hs_after_transcript_verdict(O.vp(shape), O.vp(dup), 1) == 2 | (9 << 8)        # UNSUPPORTED / stage unsupported
    near = raw.copy()
    near[11] = near[3] ^ 1                  # neighbours (same pair) are fine
    assert hostsim.hs_after_transcript_verdict(O.vp(shape), O.vp(near), 1) == 0
    assert hostsim.hs_after_transcript_verdict(O.vp(shape), O.vp(dup), 0) == 1 | (2 << 8)        # a failed PoW is reported first


@pytest.mark.parametrize("name", FIXTURES)
def test_tree_rebuild_record_equals_path_kernels(hostsim, orc, name):
    """The permutation record (output state of every transcript / per-query path permutation, what the circuit's tape evaluation
    consumes) taken from the cooperative tree rebuilds -- every node of the partial tree hashed ONCE -- is word for word the
    record the per-query path stages produce by hashing each path again from its hints, for every fixture; the per-query
    roots equal the oracle's."""
    buf, n = O.load_proof(name)
    shape = shape_of(buf)
    inputs = O.inputs_for(name)
    o = O.verify_proof(buf, n, inputs)
    hostsim.hs_verify_batch.restype = ctypes.c_void_p
    hostsim.hs_perm_record.restype = ctypes.c_uint32
    words, off = pack([(buf, n)])
    idx, vals = np.array(inputs[0], dtype=np.uint32), np.array(inputs[1], dtype=np.uint32)
    recs = {}
    for mode in (3, 7, 1):               # coop trees produce the record | coop trees + path stages | thread-per-tree + path stages
        dt = (Detail * 1)()
        ws = np.zeros(4096, dtype=np.uint8)
        base = hostsim.hs_verify_batch(O.vp(words), O.vp(off), 1, O.vp(shape), O.vp(idx), O.vp(vals), idx.size, mode, dt, O.vp(ws))
        total = hostsim.hs_perm_record(O.vp(ws), 0, None, None)
        rec = np.zeros((total, 16), dtype=np.uint32)
        rec_in = np.zeros((total, 16), dtype=np.uint32)
        trees = ctypes.c_uint32(0)
        hostsim.hs_perm_record_inputs_to(O.vp(rec_in))
        hostsim.hs_perm_record(O.vp(ws), 0, O.vp(rec), ctypes.byref(trees))
        hostsim.hs_perm_record_inputs_to(None)
        hostsim.hs_free(ctypes.c_void_p(base))
        assert dt[0].verdict == 0 and dt[0].n_perms_paths == o.n_perms_paths and dt[0].n_perms_hints == o.n_perms_hints
        assert trees.value == 5 + o.n_inner
        recs[mode] = (rec, rec_in)
    used = np.r_[np.arange(o.n_transcript_perms), np.arange(512, recs[3][0].shape[0])]     # transcript slots beyond the chain are unused
    for k in (0, 1):                                                                         # outputs, then the parallel input record
        assert np.array_equal(recs[3][k][used], recs[7][k][used]) and np.array_equal(recs[7][k][used], recs[1][k][used])
    assert recs[3][0][512:].any(axis=1).all()                                                # every path slot was written
    # the input record is what check_poseidon_invocations compares a flow entry with: output == permute(input), slot by slot
    st = np.ascontiguousarray(recs[3][1][used])
    hostsim.hs_poseidon2_permute(O.vp(st), ctypes.c_size_t(st.shape[0]))
    assert np.array_equal(st, recs[3][0][used])


@pytest.mark.parametrize("name", ["small_proof.bin", "level1-5.bin", "level7-1.bin"])
def test_build_group_strided_over_lanes(hostsim, orc, name):
    """fri::build_group_coop (k_group_coop: the alpha chain of the sample-batch coefficients strided over 16 lanes in steps of
    after_coeff^16) == the sequential fri::build_group, for 1, 2, 4, 16 and 32 emulated lanes"""
    buf, n = O.load_proof(name)
    shape = shape_of(buf)
    inputs = O.inputs_for(name)
    hostsim.hs_verify_batch.restype = ctypes.c_void_p
    words, off = pack([(buf, n)])
    idx, vals = np.array(inputs[0], dtype=np.uint32), np.array(inputs[1], dtype=np.uint32)
    dt = (Detail * 1)()
    ws = np.zeros(4096, dtype=np.uint8)
    base = hostsim.hs_verify_batch(O.vp(words), O.vp(off), 1, O.vp(shape), O.vp(idx), O.vp(vals), idx.size, 3, dt, O.vp(ws))
    assert dt[0].verdict == 0
    for G in (1, 2, 4, 16, 32):
        assert hostsim.hs_build_group_lanes(O.vp(ws), G) == 0, G
    hostsim.hs_free(ctypes.c_void_p(base))
