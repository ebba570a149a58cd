#!/usr/bin/env python3
"""Times the batched verifier on replicas of a fixture (device-resident and host-entry), both modes."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "small_proof.bin"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
blob = open(os.path.join(ROOT, "tests", "golden", "proofs", name), "rb").read()
inputs = pkg.INPUTS_SINGLE if name.startswith("small") else pkg.INPUTS_RECURSIVE
vb = pkg.VerifyBatch([blob] * n, inputs=inputs)
res = {"fixture": name, "n_proofs": n, "perms_per_proof_paths": pkg.proof_perms(vb.shape)}
for full in (True, False):
    for _ in range(2):
        v, s = vb.run(full=full)
    torch.cuda.synchronize()
    assert int(v.sum().item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        vb.run(full=full)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["full" if full else "verdict_only"] = {"ms": ms, "proofs_per_s": n / ms * 1e3}
dt = vb.fetch(0, "detail")
res["perms_hints"], res["perms_paths"] = dt.n_perms_hints, dt.n_perms_paths
t0 = time.perf_counter()
v, s = pkg.verify_proofs([blob] * n, inputs=inputs, full=True)
res["host_entry_full"] = {"ms": (time.perf_counter() - t0) * 1e3}
assert not v.any()
print(json.dumps(res))
