"""Composition of the levelised tape of a fixture's verifier circuit (host only): per level the bundles, instructions, permutations and the
longest bundle -- what the per-level clocks of tools/level_clock.py are read against.  python tools/tape_profile.py [fixture]"""
import ctypes, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O  # noqa: E402  (fixture loading only)
from verify_common import shape_of  # noqa: E402
out = os.path.join(ROOT, "build", "libhostsim.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
hs = ctypes.CDLL(out)
hs.hs_circuit_record.restype = ctypes.c_void_p
name = sys.argv[1] if len(sys.argv) > 1 else "small_proof.bin"
buf, n = O.load_proof(name)
shape = shape_of(buf)
inputs = O.inputs_for(name)
idx, vals = np.array(inputs[0], dtype=np.uint32), np.array(inputs[1], dtype=np.uint32)
h = ctypes.c_void_p(hs.hs_circuit_record(O.vp(shape), O.vp(idx), O.vp(vals), idx.size, 1))
info = np.zeros(10, dtype=np.uint32)
hs.hs_circuit_info(h, O.vp(info))
L = int(info[7])
prof = np.zeros(5 * L, dtype=np.uint32)
hs.hs_circuit_level_profile(h, O.vp(prof))
st = np.zeros(8, dtype=np.uint32)
hs.hs_circuit_bundle_stats(h, O.vp(st))
print(json.dumps({"fixture": name, "levels": L, "bundles": int(st[0]), "instructions": int(st[1]), "not_first_in_bundle": int(st[2]),
                  "operand_from_instruction_before": int(st[3]), "columns": ["bundles", "instructions", "permutations", "max_perms_in_bundle", "max_bundle_len"],
                  "per_level": prof.reshape(L, 5).tolist()}))
