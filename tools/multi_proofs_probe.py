"""Where the 256-proof mixed batch (BASELINE configs[3]) spends its time: per shape group, device-resident verification vs the host entry."""
import importlib, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
d = os.path.join(ROOT, "tests", "golden", "proofs")
names = bench.fixture_names()
blobs = {f: open(os.path.join(d, f), "rb").read() for f in names}
order = [names[i % 15] for i in range(256)]
groups = {}
for f in order:
    groups.setdefault(tuple(pkg.proof_shape(blobs[f]).key()), []).append(f)
out = []
for k, fs in groups.items():
    inputs = pkg.INPUTS_SINGLE if fs[0].startswith("small") else pkg.INPUTS_RECURSIVE
    bl = [blobs[f] for f in fs]
    vb = pkg.VerifyBatch(bl, inputs=inputs)
    for _ in range(2):
        vb.run(full=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        vb.run(full=True)
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) / 5 * 1e3
    vb.run(full=True, timed=True)
    torch.cuda.synchronize()
    st = {a: round(b, 3) for a, b in vb.stage_ms().items()}
    for _ in range(2):
        pkg.verify_proofs(bl, inputs=inputs)
    t0 = time.perf_counter()
    for _ in range(5):
        pkg.verify_proofs(bl, inputs=inputs)
    host_ms = (time.perf_counter() - t0) / 5 * 1e3
    out.append({"shape": k, "n": len(fs), "bytes": sum(len(b) for b in bl), "dev_ms": round(dev_ms, 3), "host_ms": round(host_ms, 3), "stage_ms": st})
    print(json.dumps(out[-1]))
    del vb
rest = [blobs[f] for f in order if not f.startswith("small")]
for _ in range(2):
    pkg.verify_proofs(rest)
t0 = time.perf_counter()
for _ in range(5):
    pkg.verify_proofs(rest)
print("all recursive fixtures in one call: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
