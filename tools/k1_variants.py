#!/usr/bin/env python3
"""Times the K1 permutation kernel in both code shapes (rolled / unrolled) on 2^22 states (256 MB > L2)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
st = torch.randint(0, 2**31 - 1, (n, 16), dtype=torch.int32, device="cuda")
res = {}
for variant, name in ((0, "rolled"), (1, "unrolled"), (2, "rolled_x2"), (3, "rolled_occ8"), (4, "rolled_occ10"), (5, "rolled_occ5"),
                      (10, "rolled_2_blocks_per_sm"), (12, "rolled_x2_2_blocks_per_sm"), (20, "rolled_3_blocks_per_sm"), (22, "rolled_x2_3_blocks_per_sm")):
    for _ in range(3):
        pkg.poseidon2_permute(st, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        pkg.poseidon2_permute(st, variant=variant)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[name] = {"ms": ms, "gperms_per_s": n / ms / 1e6}
print(json.dumps({"n_states": n, **res}))
