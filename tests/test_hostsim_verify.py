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
