"""Per-level time of the grid-wide tape evaluation (K6): python tools/level_clock.py [--proofs N] [--fixture F]"""
import argparse, importlib, json, os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--proofs", type=int, default=4096)
ap.add_argument("--fixture", default="small_proof.bin")
ap.add_argument("--last-layer", action="store_true")
args = ap.parse_args()
pkg = importlib.import_module("recursive-stwo_b200")
_lib = importlib.import_module("recursive-stwo_b200._lib")
pkg.init(0)
blob = open(os.path.join(ROOT, "tests", "golden", "proofs", args.fixture), "rb").read()
inputs = pkg.INPUTS_SINGLE if args.fixture.startswith("small") else pkg.INPUTS_RECURSIVE
vb = pkg.VerifyBatch([blob] * args.proofs, inputs=inputs)
vb.run(full=True)
circ = pkg.VerifierCircuit(vb.shape, inputs=inputs, last_layer=args.last_layer)
L = circ.info.n_levels
for _ in range(2):
    circ.trace(vb, check=False, export=False, preprocessed=False)
ts = torch.zeros(L + 1, dtype=torch.int64, device="cuda")
_lib.call("stwo_b200_cs_eval_level_clock", ctypes.c_void_p(ts.data_ptr()))
circ.trace(vb, check=False, export=False, preprocessed=False)
torch.cuda.synchronize()
_lib.call("stwo_b200_cs_eval_level_clock", ctypes.c_void_p(0))
t = ts.cpu().numpy()
d = np.diff(t) / 1e3
print(json.dumps({"proofs": args.proofs, "levels": L, "total_us": float(d.sum()), "level_us": [round(float(x), 2) for x in d]}))
