"""GPU tier: the batched verifier through the C ABI, every intermediate value bit-exact against the oracle."""
import ctypes
import os

import numpy as np
import pytest

import oracle_py as O
from verify_common import compare_arrays, compare_detail

pytestmark = pytest.mark.gpu
FIXTURES = sorted(f for f in os.listdir(O.PROOFS_DIR) if f.endswith(".bin") and f != "level14-1.bin")
REGIONS = ["commitment0", "sampled0", "pow_nonce", "last_coeffs", "queried0", "hash_witness0", "queried3", "fri_first_witness",
           "fri_first_hash_witness", "fri_inner0_witness", "fri_inner0_hash_witness", "fri_inner_last_witness"]


def blob_bytes(name):
    buf, n = O.load_proof(name)
    return bytes(buf[:n])


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_every_value_matches_oracle(pkg, gpu, orc, name):
    buf, n = O.load_proof(name)
    inputs = O.inputs_for(name)
    o = O.verify_proof(buf, n, inputs)
    vb = pkg.VerifyBatch([bytes(buf[:n])] * 3, inputs=inputs)      # 3 replicas: lanes must not interfere
    verdict, stage = vb.run(full=True)
    assert verdict.cpu().numpy().tolist() == [0, 0, 0] and stage.cpu().numpy().tolist() == [0, 0, 0]
    for p in (0, 2):
        dt = vb.fetch(p, "detail")
        compare_detail(dt, o)
        assert dt.n_perms_hints == o.n_perms_hints
        arrs = {k: vb.fetch(p, k) for k in ("domain_points", "answers", "circle_folds", "line_folds", "last_evals", "path_roots")}
        compare_arrays(arrs, o, o.n_queries)


@pytest.mark.parametrize("name", ["small_proof.bin", "level13-1.bin", "level1-5.bin"])
def test_tampered_batch(pkg, gpu, orc, name):
    buf, n = O.load_proof(name)
    inputs = O.inputs_for(name)
    offs = O.proof_offsets(buf, n)
    blobs, raw = [bytes(buf[:n])], [(buf, n)]
    for r in REGIONS:
        bad = buf.copy()
        bad[offs[r]] ^= 1
        blobs.append(bytes(bad[:n]))
        raw.append((bad, n))
    blobs.append(bytes(buf[: n - 4]))
    raw.append((buf, n - 4))
    vb = pkg.VerifyBatch(blobs, inputs=inputs)
    verdict, stage = vb.run(full=True)
    verdict, stage = verdict.cpu().numpy(), stage.cpu().numpy()
    seen = set()
    for p, (b, ln) in enumerate(raw):
        o = O.verify_proof(b, ln, inputs)
        assert (verdict[p], stage[p]) == (o.verdict, o.stage), (p, verdict[p], stage[p], o.verdict, o.stage)
        compare_detail(vb.fetch(p, "detail"), o, full=False)
        seen.add(O.STAGES[o.stage])
    assert {"ok", "parse", "pow", "merkle", "fri_first", "fri_inner"} <= seen
    # verdict-only mode gives the same verdicts without the per-query recomputation
    v2, s2 = vb.run(full=False)
    assert np.array_equal(v2.cpu().numpy(), verdict) and np.array_equal(s2.cpu().numpy(), stage)


def test_host_entry_mixed_shapes(pkg, gpu, orc):
    """BASELINE configs[3] in miniature: the Poseidon31 fixtures cycled, one bit-flipped copy of each, plus junk."""
    names = [f for f in FIXTURES if f != "small_proof.bin"]
    blobs, want = [], []
    for k in range(2 * len(names)):
        buf, n = O.load_proof(names[k % len(names)])
        if k >= len(names):
            buf = buf.copy()
            buf[O.proof_offsets(buf, n)["queried0"] + (k % 7)] ^= 1 << (k % 5)
        o = O.verify_proof(buf, n, O.INPUTS_RECURSIVE)
        blobs.append(bytes(buf[:n]))
        want.append((o.verdict, o.stage))
    blobs += [b"", b"\x01\x02\x03\x04" * 100]
    want += [(1, 1), (1, 1)]
    verdict, stage = pkg.verify_proofs(blobs, inputs=pkg.INPUTS_RECURSIVE, full=True)
    assert list(zip(verdict.tolist(), stage.tolist())) == want
    assert sum(1 for v, _ in want if v == 0) == len(names)


def test_small_proof_inputs(pkg, gpu, orc):
    b = blob_bytes("small_proof.bin")
    v, s = pkg.verify_proofs([b, b], inputs=pkg.INPUTS_SINGLE)
    assert v.tolist() == [0, 0]
    v, s = pkg.verify_proofs([b], inputs=pkg.INPUTS_RECURSIVE)
    assert (v[0], pkg.STAGES[int(s[0])]) == (1, "logup")


def test_large_replica_batch_is_uniform(pkg, gpu, orc):
    """size-independent property at bench scale: 1024 replicas with every 16th tampered -> exactly those reject"""
    buf, n = O.load_proof("small_proof.bin")
    off = O.proof_offsets(buf, n)["queried0"]
    good, bad = bytes(buf[:n]), None
    t = buf.copy()
    t[off] ^= 1
    bad = bytes(t[:n])
    blobs = [bad if p % 16 == 5 else good for p in range(1024)]
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    verdict, stage = vb.run(full=True)
    verdict = verdict.cpu().numpy()
    assert np.array_equal(np.nonzero(verdict)[0], np.arange(5, 1024, 16))
    assert (stage.cpu().numpy()[verdict != 0] == 5).all()


def test_pinned_host_entry_matches_device_entry(pkg, gpu, orc):
    """stwo_b200_verify_proofs_batch_pinned_dev: per-slice upload inside the call (sliced path: >= 256 proofs, tampered mixed in)"""
    buf, n = O.load_proof("small_proof.bin")
    offs = O.proof_offsets(buf, n)
    blobs = []
    for k in range(300):
        b = buf.copy()
        if k % 37 == 5:
            b[offs["queried0"] + k % 64] ^= 2
        blobs.append(bytes(b[:n]))
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    v1, s1 = vb.run(full=True)
    v1, s1 = v1.cpu().numpy().copy(), s1.cpu().numpy().copy()
    vb.d_words.zero_()                                   # the device copy must come from the pinned host buffer
    v2, s2 = vb.run_from_host(full=True)
    assert np.array_equal(v1, v2.cpu().numpy()) and np.array_equal(s1, s2.cpu().numpy())
    assert v1.sum() == sum(1 for k in range(300) if k % 37 == 5)


def test_verify_stream_double_buffer(pkg, gpu, orc):
    """VerifyStream: batches fed while the previous one is verified come out with their own verdicts, in order"""
    import torch
    buf, n = O.load_proof("small_proof.bin")
    offs = O.proof_offsets(buf, n)
    good = bytes(buf[:n])

    def batch(seed):
        out = []
        for k in range(64):
            b = buf.copy()
            if (k + seed) % 7 == 0:
                b[offs["queried0"] + (k + seed) % 64] ^= 2
            out.append(bytes(b[:n]))
        return out

    vs = pkg.VerifyStream([good] * 64, inputs=pkg.INPUTS_SINGLE)
    batches = [batch(s) for s in range(5)]
    vs.feed(batches[0])
    got = []
    for k in range(5):
        if k + 1 < 5:
            vs.feed(batches[k + 1])
        bt = vs.take()
        v, s = bt.run(full=True)
        got.append(v.clone())                    # device-side copy, queued: the host must NOT wait here (the CPU runs ahead of the GPU)
        vs.release(bt)
    torch.cuda.synchronize()
    for k in range(5):
        want = np.array([1 if (j + k) % 7 == 0 else 0 for j in range(64)], dtype=np.uint8)
        assert np.array_equal(got[k].cpu().numpy() != 0, want != 0), k


def test_weak_config_forgery_is_rejected(pkg, gpu, orc):
    """The PcsConfig is the caller's (components/hints/src/fiat_shamir.rs:69-73).  small_proof.bin with pow_bits rewritten to 0 is a
    proof that would be ACCEPTED under its own header (the oracle without a caller config accepts it): every entry point rejects it."""
    import ctypes
    buf, n = O.load_proof("small_proof.bin")
    good = bytes(buf[:n])
    f = buf.copy()
    f.view(np.uint32)[10] = 0                            # pow_bits
    forged = bytes(f[:n])
    assert O.verify_proof(f, n, O.INPUTS_SMALL).verdict == 0
    assert O.STAGES[O.verify_proof(f, n, O.INPUTS_SMALL, config=pkg.CONFIG_SINGLE.key()).stage] == "parse"
    # host entry, default allow-list (the reference's seven configs) and a one-config list that does not include the proof's
    v, s = pkg.verify_proofs([good, forged, good], inputs=pkg.INPUTS_SINGLE)
    assert v.tolist() == [0, 1, 0] and pkg.STAGES[int(s[1])] == "parse"
    v, s = pkg.verify_proofs([good, forged], inputs=pkg.INPUTS_SINGLE, config=pkg.CONFIG_STANDARD)
    assert v.tolist() == [1, 1] and [pkg.STAGES[int(x)] for x in s] == ["parse", "parse"]
    v, s = pkg.verify_proofs([good, forged], inputs=pkg.INPUTS_SINGLE, config=pkg.CONFIG_SINGLE)
    assert v.tolist() == [0, 1]
    with pytest.raises(pkg.StwoB200Error):               # no config at all is a caller error
        pkg.verify_proofs([good], inputs=pkg.INPUTS_SINGLE, config=[])
    # device entry: the shape is the caller's; a forged blob inside the batch fails Shape::matches
    with pytest.raises(ValueError):
        pkg.VerifyBatch([forged], inputs=pkg.INPUTS_SINGLE)
    vb = pkg.VerifyBatch([good, forged, good], inputs=pkg.INPUTS_SINGLE)
    assert vb.shape.key() == pkg.shape_from_config(pkg.CONFIG_SINGLE, 4, 8).key()
    v, s = vb.run(full=True)
    assert v.cpu().numpy().tolist() == [0, 1, 0] and pkg.STAGES[int(s.cpu().numpy()[1])] == "parse"
    # more queries / another blow-up claimed in the header: same
    for word, val in ((13, 1), (11, 1), (12, 1)):
        g = buf.copy()
        g.view(np.uint32)[word] = val
        v, s = pkg.verify_proofs([bytes(g[:n])], inputs=pkg.INPUTS_SINGLE)
        assert (v[0], pkg.STAGES[int(s[0])]) == (1, "parse")


def test_garbage_header_words(pkg, gpu, orc):
    """0xFFFFFFFF header fields must neither wrap into a plausible shape nor stall the batch (a wrapped log_size_plonk would send
    the OODS stage into a 2^32-step loop): parse rejects, the neighbours are verified, and the call returns promptly"""
    import time
    buf, n = O.load_proof("small_proof.bin")
    blobs = [bytes(buf[:n])]
    for word in (0, 1, 11):
        for val in (0xFFFFFFFF, 0xFFFFFFFB, 0x80000004, 29, 0):
            g = buf.copy()
            g.view(np.uint32)[word] = val
            blobs.append(bytes(g[:n]))
    vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
    t0 = time.time()
    v, s = vb.run(full=True)
    v, s = v.cpu().numpy(), s.cpu().numpy()
    assert time.time() - t0 < 5.0
    assert v.tolist() == [0] + [1] * (len(blobs) - 1) and (s[1:] == 1).all()
    v2, s2 = pkg.verify_proofs(blobs, inputs=pkg.INPUTS_SINGLE)
    assert np.array_equal(v2, v) and np.array_equal(s2, s)


@pytest.mark.parametrize("name,n", [("small_proof.bin", 300), ("recursive_proof_16_15.bin", 40), ("level1-5.bin", 5), ("level13-1.bin", 33)])
def test_tree_rebuild_record_equals_path_kernels(pkg, gpu, orc, name, n):
    """Full mode takes the circuit's permutation record from the cooperative tree rebuilds (every node hashed once).  The
    thread-per-path kernels (STWO_B200_VERIFY_PATH_KERNELS: every path hashed again from its hints) are the checker: same
    record word for word, same per-query roots, same counters -- on the sliced (300 proofs) and the one-stream launch path."""
    buf, ln = O.load_proof(name)
    inputs = O.inputs_for(name)
    o = O.verify_proof(buf, ln, inputs)
    vb = pkg.VerifyBatch([bytes(buf[:ln])] * n, inputs=inputs)
    got = {}
    for pk in (False, True):
        v, _ = vb.run(full=True, path_kernels=pk)
        assert not v.cpu().numpy().any()
        got[pk] = [(vb.fetch(p, "perm_record").copy(), vb.fetch(p, "path_roots").copy(), vb.fetch(p, "detail").n_perms_paths,
                    int(vb.fetch(p, "record_trees")[0]), vb.fetch(p, "perm_record_inputs").copy()) for p in (0, n // 2, n - 1)]
    used = np.r_[np.arange(o.n_transcript_perms), np.arange(512, got[True][0][0].shape[0])]
    for a, b in zip(got[False], got[True]):
        assert np.array_equal(a[0][used], b[0][used]) and np.array_equal(a[1], b[1])
        assert a[2] == b[2] == o.n_perms_paths and a[3] == b[3] == 5 + o.n_inner
        assert a[0][512:].any(axis=1).all()
        assert np.array_equal(a[4][used], b[4][used])                      # the parallel INPUT record, from both producers
    # ... and it is the input of the recorded output: permute(input) == output, slot by slot (the oracle's permutation as the checker)
    st = np.ascontiguousarray(got[False][0][4][used])
    orc.orc_poseidon2_permute_batch(O.vp(st), ctypes.c_size_t(st.shape[0]))
    assert np.array_equal(st, got[False][0][0][used])
    # a verdict-only run leaves no usable record: the marker is reset, the tape evaluation then permutes itself
    vb.run(full=False)
    assert int(vb.fetch(0, "record_trees")[0]) == 0
