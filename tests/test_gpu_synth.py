"""GPU tier: the synthetic FRI + Merkle instances of BASELINE configs[4] part i, generated on the device and verified through the C ABI.
Checker: the oracle's FRI-only verifier (orc_fri_verify_synth) on a sample of >= 64 instances -- draws, folds, per-query roots and
permutation counts bit-exact -- and the host build of the same seeds (tests/hostsim), which must produce the very same blobs."""
import ctypes

import numpy as np
import pytest

import oracle_py as O
from test_hostsim_synth import generate, oracle_verify

pytestmark = pytest.mark.gpu


def _shape_S(pkg):
    return pkg.shape_from_config(pkg.PcsConfig(0, 5, 2, 16), 4, 8)          # small_proof.bin's shape, no proof of work


def test_device_generator_equals_host_generator(pkg, gpu, hostsim, orc):
    for cfg, lsp, lspos in ((pkg.PcsConfig(3, 2, 1, 5), 3, 4), (pkg.PcsConfig(0, 1, 2, 7), 5, 4), (pkg.PcsConfig(2, 3, 0, 9), 2, 6)):
        shape = pkg.shape_from_config(cfg, lsp, lspos)
        sb = pkg.SynthBatch(shape, 9, seed0=100)
        for p in (0, 4, 8):
            want = generate(hostsim, shape.key(), 100 + p)
            assert np.array_equal(sb.blob(p), want), (shape.key(), p)


def test_sampled_parity_against_the_oracle(pkg, gpu, orc):
    shape = _shape_S(pkg)
    n = 200
    sb = pkg.SynthBatch(shape, n, seed0=0, chunk=64)
    verdict, stage = sb.run(full=True)
    assert not verdict.cpu().numpy().any() and not stage.cpu().numpy().any()
    nq, n_inner = shape.n_queries, shape.n_inner
    positions = set()
    for p in list(range(0, n, 3))[:67]:
        o = oracle_verify(orc, sb.blob(p))
        assert (o.verdict, o.stage) == (0, 0)
        dt = sb.fetch(p, "detail")
        assert list(dt.fs.raw_queries)[:nq] == list(o.raw_queries)[:nq]
        for i in range(n_inner + 1):
            assert list(dt.fs.fri_alphas[i]) == list(o.fri_alphas[i])
        assert dt.n_perms_hints == o.n_perms_hints and dt.n_perms_paths == o.n_perms_paths
        for g in range(o.n_logs):
            assert np.array_equal(sb.fetch(p, "circle_folds")[g, :nq], np.ctypeslib.as_array(o.circle_folds)[g, :nq])
        lf = sb.fetch(p, "line_folds")
        for li in range(n_inner):
            assert np.array_equal(lf[li, :nq], np.ctypeslib.as_array(o.line_folds)[li, :nq])
        assert np.array_equal(sb.fetch(p, "last_evals")[:nq], np.ctypeslib.as_array(o.last_layer_evals)[:nq])
        roots = sb.fetch(p, "path_roots")
        want = np.ctypeslib.as_array(o.path_roots)
        for t in range(4, 5 + n_inner):
            assert np.array_equal(roots[t, :nq], want[t, :nq]), t
        positions.add(tuple(o.raw_queries)[:nq])
    assert len(positions) == 67                                  # every instance opens its own positions
    # the record of the FRI trees from the rebuilds == from the thread-per-path kernels
    rec = sb.fetch(7, "perm_record").copy()
    sb.run(full=True, path_kernels=True)
    rec2 = sb.fetch(7, "perm_record")
    n_fri = int(o.n_perms_paths - o.n_transcript_perms)          # the FRI trees' slots are the tail of the record (verify.cuh HintLayout); the
    assert n_fri == 2112                                         # commitment-tree slots of a real proof stay unused for an instance
    assert rec[-n_fri:].any(axis=1).all() and np.array_equal(rec[-n_fri:], rec2[-n_fri:])


def test_tampered_instances(pkg, gpu, orc):
    import torch
    shape = _shape_S(pkg)
    sb = pkg.SynthBatch(shape, 40, seed0=900)
    hdr = sb.blob(0)[:256]
    spots = [int(hdr[80]), int(hdr[81]), int(hdr[82]) + 5, int(hdr[84]) + 3, int(hdr[96 + 2]), int(hdr[160 + 4]) + 1, int(hdr[128 + 3]) if hdr[10 + 3] else int(hdr[82])]
    for k, off in enumerate(spots):
        sb.d_words[(3 + 4 * k) * sb.words + off] ^= 1
    verdict, stage = sb.run(full=True)
    verdict, stage = verdict.cpu().numpy(), stage.cpu().numpy()
    for p in range(40):
        o = oracle_verify(orc, sb.blob(p))
        assert (verdict[p], stage[p]) == (o.verdict, o.stage), p
    assert verdict.sum() == len(spots)


def test_folding_circuit_trace_of_instances(pkg, gpu, orc):
    """the tape of configs[4] part i: the folding-stage circuit traced for a batch of distinct instances -- variables, Poseidon flow and
    the 13 value columns against the oracle DSL's `folding` on sampled instances; every instance satisfies it; a tampered one does not"""
    from circuit_common import D
    shape = _shape_S(pkg)
    n = 70
    sb = pkg.SynthBatch(shape, n, seed0=300)
    sb.d_words[41 * sb.words + int(sb.blob(0)[82]) + 2] ^= 1                # an answer word of instance 41
    verdict, _ = sb.run(full=True)
    verdict = verdict.cpu().numpy()
    assert verdict[41] == 1 and verdict.sum() == 1
    circ = pkg.VerifierCircuit(shape, folding=True)
    assert circ.info.n_flow == 2112 and circ.info.num_input == 3        # no public inputs beyond the three constants
    r = circ.trace(sb, check=True, export=True)
    bad_row, bad_flow = r["bad_row"].cpu().numpy(), r["bad_flow"].cpu().numpy()
    assert bad_row[41] >= 0 and (np.delete(bad_row, 41) == -1).all() and (bad_flow == -1).all()
    for p in (0, 33, n - 1):
        cs, _ = D.folding_circuit(sb.blob(p), O.VerifyOut)
        assert np.array_equal(circ.fetch(p, "variables"), np.array(cs.variables, dtype=np.uint32))
        wire, addr, wh, wsw = cs.flow_arrays()
        assert np.array_equal(circ.fetch(p, "flow_hash"), wh) and np.array_equal(circ.fetch(p, "flow_swap"), wsw)
        got = pkg.VerifierCircuit.assemble_trace(r["preprocessed"], r["values"][p])
        assert np.array_equal(got, cs.trace_columns())
    # the same trace with every permutation recomputed by the tape instead of taken from the tree rebuilds' record
    keep = r["values"].clone()
    r2 = circ.trace(sb, check=True, export=True, native_hints=False)
    assert (r2["values"] == keep).all()


def test_folding_circuit_on_a_real_proof(pkg, gpu, orc):
    """the same circuit over the workspace of a real proof's verification = the folding part of its verifier circuit"""
    import os
    blob = open(os.path.join(O.PROOFS_DIR, "small_proof.bin"), "rb").read()
    vb = pkg.VerifyBatch([blob] * 33, inputs=pkg.INPUTS_SINGLE)
    verdict, _ = vb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(vb.shape, folding=True)
    r = circ.trace(vb, check=True, export=True)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    assert (r["values"][0] == r["values"][32]).all()


def test_folding_circuit_full_batch(pkg, gpu, orc):
    """BASELINE configs[4] part i at full size: 4096 distinct instances, every one accepted and its folding circuit satisfied"""
    shape = _shape_S(pkg)
    sb = pkg.SynthBatch(shape, 4096, seed0=5000)
    verdict, _ = sb.run(full=True)
    assert not verdict.cpu().numpy().any()
    circ = pkg.VerifierCircuit(shape, folding=True)
    r = circ.trace(sb, check=True, export=True, preprocessed=False)
    assert (r["bad_row"].cpu().numpy() == -1).all() and (r["bad_flow"].cpu().numpy() == -1).all()
    v = r["values"]
    assert not (v[0] == v[1]).all() and not (v[100] == v[4095]).all()      # distinct instances, distinct traces
