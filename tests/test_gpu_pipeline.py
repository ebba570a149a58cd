"""GPU tier: VerifyTracePipeline (upload | verification | trace pass of neighbouring batches on their own streams, three slots): every
batch of a stream comes out with its own verdicts and check results, in order, and the trace columns equal the unpipelined ones."""
import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu


def test_pipeline_results_per_batch(pkg, gpu, orc):
    import torch
    buf, n = O.load_proof("small_proof.bin")
    offs = O.proof_offsets(buf, n)
    good = bytes(buf[:n])

    def batch(seed):
        out = []
        for k in range(96):
            b = buf.copy()
            if (k + seed) % 5 == 0:
                b[offs["queried0"] + (k + seed) % 64] ^= 2          # rejected at the Merkle stage
            if (k + 2 * seed) % 11 == 3:
                b[offs["sampled0"] + 5] ^= 1                          # rejected at the OODS stage
            out.append(bytes(b[:n]))
        return out

    pipe = pkg.VerifyTracePipeline([good] * 96, inputs=pkg.INPUTS_SINGLE, n_slots=3)
    batches = [batch(s) for s in range(7)]
    got = []
    for b in batches:                                  # the host never waits inside the loop
        h = pipe.step(b)
        got.append(h)
    # results of the last n_slots steps are still in their pinned buffers
    pipe.join()
    for k in range(4, 7):
        v, s, bad_row, bad_flow = pipe.result(got[k])
        want = [O.verify_proof(np.frombuffer(b, dtype=np.uint8).copy(), len(b), O.INPUTS_SMALL) for b in batches[k][:12]]
        assert [int(x) for x in v[:12]] == [o.verdict for o in want] and [int(x) for x in s[:12]] == [o.stage for o in want]
        rejected = np.array([(j + k) % 5 == 0 or (j + 2 * k) % 11 == 3 for j in range(96)])
        assert np.array_equal(v.numpy() != 0, rejected)
        assert np.array_equal(bad_row.numpy() != -1, rejected) and (bad_flow.numpy() == -1).all()
    # the trace columns of the last step equal an unpipelined run on the same blobs
    vb = pkg.VerifyBatch(batches[6], inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    r = circ.trace(vb, check=True, export=True, preprocessed=False)
    ok = ~torch.from_numpy(np.array([(j + 6) % 5 == 0 or (j + 12) % 11 == 3 for j in range(96)])).to(gpu)
    assert torch.equal(pipe.values[ok], r["values"][ok])


def test_pipeline_resident_replay(pkg, gpu, orc):
    """upload=False (the bench's device-resident leg): the slots' device copies are used as they are"""
    buf, n = O.load_proof("level13-1.bin")
    pipe = pkg.VerifyTracePipeline([bytes(buf[:n])] * 40, inputs=pkg.INPUTS_RECURSIVE, n_slots=2)
    for _ in range(5):
        h = pipe.step(upload=False)
    pipe.join()
    v, s, bad_row, bad_flow = pipe.result(h)
    assert not v.numpy().any() and (bad_row.numpy() == -1).all() and (bad_flow.numpy() == -1).all()
