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
