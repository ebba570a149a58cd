"""GPU tier, last file of the suite: stwo_b200_shutdown releases the library's per-device resources (staging area, worker-stream pools,
trace-pass side stream and events) and a following stwo_b200_init brings everything back -- same results before and after."""
import os

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu


def test_shutdown_then_init_again(pkg, gpu, orc):
    from importlib import import_module
    _lib = import_module("recursive-stwo_b200._lib")
    blob = open(os.path.join(O.PROOFS_DIR, "small_proof.bin"), "rb").read()
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 1
    blobs = [blob, bytes(bad)] * 150                                  # 300 proofs: the sliced path with its worker-stream pool

    def run():
        vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_SINGLE)
        v, s = vb.run(full=True)
        circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
        r = circ.trace(vb, check=True, export=True, timed=True)
        return v.cpu().numpy().copy(), s.cpu().numpy().copy(), r["bad_row"].cpu().numpy().copy(), r["values"][0].cpu().numpy().copy()

    before = run()
    assert before[0][0] == 0 and before[0][1] == 1
    import torch
    torch.cuda.synchronize()
    _lib.call("stwo_b200_shutdown")
    pkg.init(0)
    after = run()
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    v, s = pkg.verify_proofs(blobs[:4], inputs=pkg.INPUTS_SINGLE)     # the host entry re-grows its staging area
    assert v.tolist() == [0, 1, 0, 1]
