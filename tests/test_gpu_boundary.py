"""GPU tier: the stage-level entry points of the boundary and the prover-facing PoseidonFlow export (8f-3), each
against the oracle on the reference's fixtures."""
import os

import numpy as np
import pytest

import oracle_py as O
from circuit_common import oracle_circuit
from verify_common import compare_detail

pytestmark = pytest.mark.gpu


def _cfg(pkg, name):
    buf, n = O.load_proof(name)
    s = pkg.proof_shape(bytes(buf[:n]))
    return next(c for c in pkg.REFERENCE_CONFIGS if c.key() == (s.pow_bits, s.log_blowup, s.log_last, s.n_queries))


def _batch(name):
    """the fixture, a PoW-breaking, an OODS-breaking and a Merkle-breaking copy, a truncated one"""
    buf, n = O.load_proof(name)
    offs = O.proof_offsets(buf, n)
    raw = [(buf, n)]
    for region in ("pow_nonce", "sampled0", "queried0"):
        b = buf.copy()
        b[offs[region]] ^= 1
        raw.append((b, n))
    raw.append((buf, n - 4))
    return raw, [bytes(b[:ln]) for b, ln in raw]


@pytest.mark.parametrize("name", ["small_proof.bin", "recursive_proof_16_15.bin", "level1-5.bin", "level13-1.bin"])
def test_channel_replay_batch(pkg, gpu, orc, name):
    """FiatShamirHints::new for a batch (components/hints/src/fiat_shamir.rs:69-307): every draw, the digest, PoW, OODS values"""
    raw, blobs = _batch(name)
    inputs = O.inputs_for(name)
    det, verdict, stage = pkg.channel_replay(blobs, _cfg(pkg, name), inputs)
    for p, (b, ln) in enumerate(raw):
        o = O.verify_proof(b, ln, inputs)
        if o.stage in (0, 5, 6, 7, 8):                  # what fails after the OODS check is not this stage's business
            assert (verdict[p], stage[p]) == (0, 0)
            o.verdict, o.stage = 0, 0
        compare_detail(det[p], o, full=False)
        assert (verdict[p], stage[p]) == (o.verdict, o.stage)


@pytest.mark.parametrize("name", ["small_proof.bin", "level2-1.bin", "level9-1.bin"])
def test_fri_answers_and_folds_batch(pkg, gpu, orc, name):
    """AnswerHints / First+InnerLayersHints for a batch: DEEP answers, domain points, circle and line folds, last-layer evaluations"""
    buf, n = O.load_proof(name)
    inputs = O.inputs_for(name)
    o = O.verify_proof(buf, n, inputs)
    blobs = [bytes(buf[:n])] * 5
    cfg = _cfg(pkg, name)
    nq = o.n_queries
    ans, pts, verdict, stage = pkg.fri_answers(blobs, cfg, inputs)
    assert not verdict.any()
    circle, line, last, verdict, stage = pkg.fri_folds(blobs, cfg, inputs)
    assert not verdict.any()
    for p in (0, 4):
        for g in range(o.n_logs):
            assert np.array_equal(ans[p, g, :nq], np.ctypeslib.as_array(o.fri_answers)[g, :nq])
            assert np.array_equal(pts[p, g, :nq], np.ctypeslib.as_array(o.domain_points)[g, :nq])
            assert np.array_equal(circle[p, g, :nq], np.ctypeslib.as_array(o.circle_folds)[g, :nq])
        for li in range(o.n_inner):
            assert np.array_equal(line[p, li, :nq], np.ctypeslib.as_array(o.line_folds)[li, :nq])
        assert np.array_equal(last[p, :nq], np.ctypeslib.as_array(o.last_layer_evals)[:nq])
    # a proof whose FRI witness is broken is NOT rejected by the fold stage's values (its Merkle check comes later), one whose
    # sampled values are broken is rejected before the answers
    offs = O.proof_offsets(buf, n)
    b = buf.copy()
    b[offs["sampled0"]] ^= 1
    _, _, v2, s2 = pkg.fri_answers([bytes(buf[:n]), bytes(b[:n])], cfg, inputs)
    ob = O.verify_proof(b, n, inputs)
    assert (v2[0], s2[0]) == (0, 0) and (v2[1], s2[1]) == (ob.verdict, ob.stage)


def test_hash_column_capacity_batch(pkg, gpu, orc, rng):
    """hash_column_get_capacity (components/hints/src/folding.rs:77): the 8/chunk zero-padded sponge with chained capacity"""
    for n_cols in (0, 1, 4, 8, 9, 40, 60):
        cols = rng.integers(0, O.P, size=(37, n_cols), dtype=np.uint32)
        want = np.zeros((37, 8), dtype=np.uint32)
        for i in range(37):
            st = np.zeros(16, dtype=np.uint32)
            for ch in range(max(1, (n_cols + 7) // 8)):
                chunk = np.zeros(8, dtype=np.uint32)
                part = cols[i, 8 * ch: 8 * ch + 8]
                chunk[: part.size] = part
                st[:8] = chunk
                st = O.permute(st[None, :])[0]
            want[i] = st[8:]
        assert np.array_equal(pkg.hash_column_capacity(cols), want), n_cols


@pytest.mark.parametrize("name", ["small_proof.bin", "level12-1.bin"])
def test_poseidon_flow_export(pkg, gpu, orc, name):
    """the prover-facing PoseidonFlow (plonk_with_poseidon.rs:117-128): entries as recorded, padding entries (0, C1), (0, C1), (0, C2),
    (0, C3) with the caller's constants up to max(32, ceil(n / 16) * 16) (:296-321)"""
    cs, out = oracle_circuit(name, 1)
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    n = 35
    inputs = pkg.INPUTS_SINGLE if name.startswith("small") else pkg.INPUTS_RECURSIVE
    vb = pkg.VerifyBatch([blob] * n, inputs=inputs)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=inputs)
    circ.trace(vb, check=True, export=False, preprocessed=False)
    consts = np.arange(1, 25, dtype=np.uint32).reshape(3, 8) * 1000003 % O.P
    fl = circ.export_flow(consts)
    wire, addr, wh, wsw = cs.flow_arrays()
    nf, n_pad = wh.shape[0], cs.n_flow_padded
    assert fl["hash"].shape == (n, n_pad, 32) and n_pad == max(32, (nf + 15) // 16 * 16)
    h, sw = fl["hash"].cpu().numpy().view(np.uint32), fl["swap"].cpu().numpy()
    pad = np.concatenate([consts[0], consts[0], consts[1], consts[2]])
    for p in (0, 31, 32, n - 1):
        assert np.array_equal(h[p, :nf], wh) and np.array_equal(sw[p, :nf], wsw)
        assert (h[p, nf:] == pad).all() and not sw[p, nf:].any()
    assert np.array_equal(fl["wire"][:nf], wire) and not fl["wire"][nf:].any() and np.array_equal(fl["swap_addr"][:nf], addr)


def test_gather_entry_points_single_rank(pkg, gpu, orc):
    """stwo_b200_comm_* / _gather_verdicts / _gather_trace_columns with a one-rank NCCL communicator created by the library itself
    (libnccl.so.2 through dlopen); the N > 1 path runs under torchrun in tools/gpu_round_multi.sh"""
    import torch
    sharding = __import__("importlib").import_module("recursive-stwo_b200.sharding")
    comm = sharding.Comm(0, 1, gpu)
    v = torch.arange(10, dtype=torch.uint8, device=gpu)
    s = (v * 3).to(torch.uint8)
    gv, gs = comm.gather_verdicts(v, s, 10)
    assert torch.equal(gv, v) and torch.equal(gs, s)
    vals = torch.arange(10 * 13 * 64, dtype=torch.int32, device=gpu).reshape(10, 13, 64)
    assert torch.equal(comm.gather_trace_columns(vals, 10, dst=0), vals)
    comm.close()
