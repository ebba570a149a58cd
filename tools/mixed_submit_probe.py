"""Host submission time of one MixedBatch step against its device time (is the mixed batch submission-bound?)"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
d = os.path.join(ROOT, "tests", "golden", "proofs")
names = bench.fixture_names()
blobs = {f: open(os.path.join(d, f), "rb").read() for f in names}
order = [names[i % len(names)] for i in range(256)]
mb = pkg.MixedBatch([blobs[f] for f in order if not f.startswith("small")], inputs=pkg.INPUTS_RECURSIVE)
for _ in range(3):
    v, s = mb.run(trace=True, export=True)
torch.cuda.synchronize()
res = {"groups": len(mb.groups), "proofs": mb.n}
k = 6
t0 = time.perf_counter()
for _ in range(k):
    mb.run(trace=True, export=True)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
res["submit_ms_per_step"] = (t1 - t0) / k * 1e3
res["total_ms_per_step"] = (t2 - t0) / k * 1e3
# one group alone: the heaviest and the lightest
for g in (mb.groups[0], mb.groups[-1]):
    circ = g.circuit
    for _ in range(2):
        g.batch.run(full=True); circ.trace(g.batch, check=True, export=True, preprocessed=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        g.batch.run(full=True); circ.trace(g.batch, check=True, export=True, preprocessed=False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res["group_%d_proofs_%d_queries" % (len(g.ids), g.batch.shape.n_queries)] = {"submit_ms": (t1 - t0) / k * 1e3, "total_ms": (t2 - t0) / k * 1e3}
print(json.dumps(res))
