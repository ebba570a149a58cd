"""Experiment: verification of batch k+1 on one stream beside the trace pass of batch k on another (two device slots)."""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("recursive-stwo_b200")
pkg.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
blob = open(os.path.join(ROOT, "tests", "golden", "proofs", "small_proof.bin"), "rb").read()
slots = [pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_SINGLE) for _ in range(2)]
circ = pkg.VerifierCircuit(slots[0].shape, inputs=pkg.INPUTS_SINGLE)
for s in slots:
    s.run(full=True)
    r = circ.trace(s, check=True, export=True, preprocessed=False)
torch.cuda.synchronize()
assert int((r["bad_row"] != -1).sum().item()) == 0

def serial(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        s = slots[i % 2]
        s.run(full=True)
        circ.trace(s, check=True, export=True, preprocessed=False)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

def piped(k, prio):
    lo, hi = -1, 0
    sv = torch.cuda.Stream(priority=lo if prio == "verify" else hi)
    st = torch.cuda.Stream(priority=lo if prio == "trace" else hi)
    verified = [torch.cuda.Event() for _ in range(2)]
    traced = [torch.cuda.Event() for _ in range(2)]
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sv.wait_stream(cur); st.wait_stream(cur)
    for e in traced:
        e.record(st)
    for i in range(k):
        s = slots[i % 2]
        with torch.cuda.stream(sv):
            sv.wait_event(traced[i % 2])          # the slot's previous trace pass has finished with its workspace
            s.run(full=True)
            verified[i % 2].record(sv)
        with torch.cuda.stream(st):
            st.wait_event(verified[i % 2])
            r = circ.trace(s, check=True, export=True, preprocessed=False)
            traced[i % 2].record(st)
    cur.wait_stream(sv); cur.wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    assert int((r["bad_row"] != -1).sum().item()) == 0 and int((r["bad_flow"] != -1).sum().item()) == 0
    return e0.elapsed_time(e1) / k

out = {"proofs": n, "serial_ms": serial(8)}
for prio in ("none", "verify", "trace"):
    piped(4, prio)
    out["piped_ms_prio_" + prio] = piped(12, prio)
print(json.dumps(out))
