"""Stage timings of trace generation on one GPU: python tools/trace_bench.py [--proofs N] [--fixture F] [--reps R]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proofs", type=int, default=4096)
    ap.add_argument("--fixture", default="small_proof.bin")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--last-layer", action="store_true", help="the circuit of examples/last-layer instead of the recursive verifier")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("recursive-stwo_b200")
    pkg.init(0)
    blob = open(os.path.join(ROOT, "tests", "golden", "proofs", args.fixture), "rb").read()
    inputs = pkg.INPUTS_SINGLE if args.fixture.startswith("small") else pkg.INPUTS_RECURSIVE
    vb = pkg.VerifyBatch([blob] * args.proofs, inputs=inputs)
    v, _ = vb.run(full=True)
    assert int(v.sum().item()) == 0
    circ = pkg.VerifierCircuit(vb.shape, inputs=inputs, last_layer=args.last_layer)
    info = {k: getattr(circ.info, k) for k, _ in circ.info._fields_}
    acc = {}
    for r in range(args.reps + 1):
        out = circ.trace(vb, check=not args.no_check, export=True, preprocessed=True, timed=True)
        torch.cuda.synchronize()
        if r:
            for k, t in circ.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + t / args.reps
    if not args.no_check:
        assert int((out["bad_row"] != -1).sum().item()) == 0 and int((out["bad_flow"] != -1).sum().item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        vb.run(full=True)
        circ.trace(vb, check=not args.no_check, export=True, preprocessed=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    e0.record()
    for _ in range(args.reps):
        circ.trace(vb, check=not args.no_check, export=True, preprocessed=False)
    e1.record()
    torch.cuda.synchronize()
    trace_ms = e0.elapsed_time(e1) / args.reps
    n, nr = args.proofs, info["n_rows"]
    export_bytes = n * nr * (3 * 16 + 13 * 4)
    print(json.dumps({"fixture": args.fixture, "proofs": n, "info": info, "trace_stage_ms": acc, "verify_plus_trace_ms": ms, "trace_untimed_ms": trace_ms,
                      "proofs_per_sec_verify_plus_trace": n / (ms * 1e-3),
                      "export_gbs": export_bytes / (acc["export"] * 1e-3) / 1e9,
                      "eval_perms_per_sec": n * info["n_flow"] / (acc["eval"] * 1e-3), "last_layer": args.last_layer,
                      "eval_tape_instructions_per_sec": n * info["n_ins"] / (acc["eval"] * 1e-3),
                      "circuit_workspace_mb": circ.workspace_bytes(n) >> 20}))


if __name__ == "__main__":
    main()
