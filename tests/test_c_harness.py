"""The C ABI exercised without Python in between: tests/c_harness/harness.c is compiled with gcc against include/stwo_b200.h, linked
to libstwo_b200.so and run as a process.  CPU tier: the host-only entry points + every compute call fails with E_NO_DEVICE.  GPU
tier: Poseidon2 KAT, the reference's fixture through stwo_b200_verify_proofs_batch and stwo_b200_channel_replay_batch."""
import os
import subprocess

import pytest

import oracle_py as O

PKG = os.path.join(O.ROOT, "recursive-stwo_b200")


def _build():
    out = os.path.join(O.ROOT, "build", "c_harness")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-o", out, os.path.join(O.ROOT, "tests", "c_harness", "harness.c"),
                           "-L" + PKG, "-lstwo_b200", "-Wl,-rpath," + PKG])
    return out


def _run(exe):
    r = subprocess.run([exe, os.path.join(O.PROOFS_DIR, "small_proof.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_c_harness_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("the GPU tier runs the device half")
    assert _run(_build()).startswith("OK no-device")


@pytest.mark.gpu
def test_c_harness_on_device(pkg, gpu):
    assert _run(_build()).startswith("OK device")
