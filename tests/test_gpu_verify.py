"""GPU tier: the batched verifier through the C ABI, every intermediate value bit-exact against the oracle."""
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
        got.append(v.cpu().numpy().copy())
        vs.release(bt)
    torch.cuda.synchronize()
    for k in range(5):
        want = np.array([1 if (j + k) % 7 == 0 else 0 for j in range(64)], dtype=np.uint8)
        assert np.array_equal(got[k] != 0, want != 0), k
