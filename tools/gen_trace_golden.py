"""Generates tests/golden/trace_digests.json: row/flow/variable counts and the SHA-256 of the 22 trace columns the oracle's
circuit DSL produces for every step of the reference's recursion chain.  Run from the repo root:  python tools/gen_trace_golden.py"""
import json
import math
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import orc_dsl as D  # noqa: E402

CHAIN = [("small_proof.bin", 1, "recursive_proof_16_15.bin"), ("recursive_proof_16_15.bin", 5, "level1-5.bin"),
         ("level1-5.bin", 1, "level2-1.bin"), ("level2-1.bin", 1, "level3-1.bin"), ("level3-1.bin", 5, "level4-5.bin"),
         ("level4-5.bin", 1, "level5-1.bin"), ("level5-1.bin", 1, "level6-1.bin"), ("level6-1.bin", 1, "level7-1.bin"),
         ("level7-1.bin", 1, "level8-1.bin"), ("level8-1.bin", 1, "level9-1.bin"), ("level9-1.bin", 1, "level10-1.bin"),
         ("level10-1.bin", 1, "level11-1.bin"), ("level11-1.bin", 1, "level12-1.bin"), ("level12-1.bin", 1, "level13-1.bin"),
         ("level13-1.bin", 1, "level14-1.bin")]
out = []
for src, mult, nxt in CHAIN:
    blob = open(os.path.join(O.PROOFS_DIR, src), "rb").read()
    inputs = D.INPUTS_SINGLE if src.startswith("small") else D.INPUTS_RECURSIVE
    cs, vo = D.verifier_circuit(blob, inputs, mult, O.VerifyOut)
    want = struct.unpack("<II", open(os.path.join(O.PROOFS_DIR, nxt), "rb").read()[:8])
    got = (int(math.log2(len(cs.a_wire))), math.ceil(math.log2(cs.n_flow_padded * 6)))
    assert got == tuple(want), (src, got, want)
    wire, addr, h, sw = cs.flow_arrays()
    out.append({"src": src, "multipliers": mult, "next": nxt, "rows": cs.n_rows_unpadded, "flow": cs.n_flow_unpadded,
                "vars": len(cs.variables), "log_rows": got[0], "log_poseidon": got[1],
                "trace_sha256": D.trace_digest(cs.trace_columns()), "flow_hash_sha256": D.trace_digest(h),
                "flow_wire_sha256": D.trace_digest(wire)})
    print(out[-1])
last = []
for name in ["level13-1.bin", "level12-1.bin", "level10-1.bin"]:
    blob = open(os.path.join(O.PROOFS_DIR, name), "rb").read()
    cs, vo = D.last_layer_circuit(blob, O.VerifyOut)
    last.append({"src": name, "rows": cs.n_rows_unpadded, "log_rows": len(cs.a_wire).bit_length() - 1, "vars": len(cs.variables),
                 "public_inputs": cs.num_input, "emulated_permutations": cs.n_perm, "trace_sha256": D.trace_digest(cs.trace_columns())})
    print(last[-1])
json.dump({"generator": "tools/gen_trace_golden.py", "rows_per_permutation": 6, "chain": out, "last_layer": last},
          open(os.path.join(ROOT, "tests", "golden", "trace_digests.json"), "w"), indent=1)
