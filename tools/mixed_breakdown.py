"""per shape group of the configs[3] batch: proofs, verify stage ms, trace stage ms (1 GPU)"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
d = os.path.join(ROOT, "tests", "golden", "proofs")
names = bench.fixture_names()
blobs = [open(os.path.join(d, names[i % len(names)]), "rb").read() for i in range(256) if not names[i % len(names)].startswith("small")]
mb = pkg.MixedBatch(blobs, inputs=pkg.INPUTS_RECURSIVE)
mb.run(); mb.run()
for g in mb.groups:
    for _ in range(2):
        g.batch.run(full=True, timed=True); vs = g.batch.stage_ms()
        g.circuit.trace(g.batch, check=True, export=True, preprocessed=False, timed=True); ts = g.circuit.stage_ms()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.batch.run(full=True); g.circuit.trace(g.batch, check=True, export=True, preprocessed=False)
    e1.record(); torch.cuda.synchronize()
    i = g.circuit.info
    print(json.dumps({"shape": list(g.batch.shape.key()), "n": len(g.ids), "rows": i.n_rows, "flow": i.n_flow, "levels": i.n_levels, "ms": round(e0.elapsed_time(e1) / 3, 3),
                      "verify": {k: round(v, 3) for k, v in vs.items() if v > 0.02}, "trace": {k: round(v, 3) for k, v in ts.items() if v > 0.02}}))
