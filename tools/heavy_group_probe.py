"""Stage times of one small group of the heaviest shape (34 proofs of level1-5.bin: 80 queries, 2^19-row circuit), the chain that bounds the mixed batch"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
name, n = (sys.argv[1] if len(sys.argv) > 1 else "level1-5.bin"), int(sys.argv[2]) if len(sys.argv) > 2 else 34
blob = open(os.path.join(ROOT, "tests", "golden", "proofs", name), "rb").read()
vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_RECURSIVE)
circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_RECURSIVE)
acc = {}
for r in range(4):
    vb.run(full=True, timed=True)
    a = dict(vb.stage_ms())
    circ.trace(vb, check=True, export=True, preprocessed=False, timed=True)
    a.update({"trace_" + k: v for k, v in circ.stage_ms().items()})
    if r:
        for k, v in a.items():
            acc[k] = acc.get(k, 0.0) + v / 3
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    vb.run(full=True); circ.trace(vb, check=True, export=True, preprocessed=False)
e1.record(); torch.cuda.synchronize()
info = {k: getattr(circ.info, k) for k in ("n_rows", "n_vars", "n_flow", "n_ins", "n_levels")}
print(json.dumps({"fixture": name, "proofs": n, "shape": list(vb.shape.key()), "circuit": info, "stage_ms": {k: round(v, 3) for k, v in acc.items() if v > 0.01},
                  "sum_ms": round(sum(acc.values()), 2), "untimed_ms_per_step": round(e0.elapsed_time(e1) / 4, 2)}))
