"""How long the host takes to SUBMIT one pipeline step (no device wait) against how long the device takes to run it: small batches are
submission-bound when the two are close.  python tools/host_submit_probe.py [proofs]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
out = []
for n in [int(a) for a in sys.argv[1:]] or [512, 4096]:
    blob = open(os.path.join(ROOT, "tests", "golden", "proofs", "small_proof.bin"), "rb").read()
    pipe = pkg.VerifyTracePipeline([blob] * n, inputs=pkg.INPUTS_SINGLE)
    for _ in range(4):
        pipe.step(upload=False)
    pipe.join()
    res = {"proofs": n}
    for upload in (False, True):
        k = 24
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            pipe.step(upload=upload)
        t1 = time.perf_counter()
        pipe.join()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res["upload" if upload else "resident"] = {"submit_ms_per_step": (t1 - t0) / k * 1e3, "total_ms_per_step": (t2 - t0) / k * 1e3}
    # the two entry points alone, on one stream, host time only
    vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_SINGLE)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_SINGLE)
    vb.run(full=True); circ.trace(vb, check=True, export=True, preprocessed=False); torch.cuda.synchronize()
    for name, fn in (("verify_entry", lambda: vb.run(full=True)), ("trace_entry", lambda: circ.trace(vb, check=True, export=True, preprocessed=False))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res[name] = {"submit_ms": (t1 - t0) / 20 * 1e3, "total_ms": (t2 - t0) / 20 * 1e3}
    out.append(res)
    del pipe, vb, circ
    torch.cuda.empty_cache()
print(json.dumps(out))
