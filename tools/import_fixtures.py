#!/usr/bin/env python3
"""Imports the reference's proof fixtures (test DATA, not source) into tests/golden/proofs/ so that the
GPU box — where /root/reference does not exist — can run the parity tests.  The 15 Poseidon31-hash proofs are
the golden inputs of the reference's own Pattern-C tests and examples:
  components/test_data/{small_proof,recursive_proof_16_15}.bin, examples/multi-proofs/data/level{1..13}-*.bin
(level14-1 == hybrid_hash.bin uses the SHA-256 hybrid hasher and is imported only as a must-not-parse case)."""
import glob, hashlib, json, os, shutil, sys
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "tests", "golden", "proofs")
os.makedirs(dst, exist_ok=True)
files = [os.path.join(REF, "components/test_data/small_proof.bin"), os.path.join(REF, "components/test_data/recursive_proof_16_15.bin")]
files += sorted(glob.glob(os.path.join(REF, "examples/multi-proofs/data/level*.bin")))
manifest = {}
for f in files:
    name = os.path.basename(f)
    shutil.copyfile(f, os.path.join(dst, name))
    manifest[name] = {"source": os.path.relpath(f, REF), "bytes": os.path.getsize(f), "sha256": hashlib.sha256(open(f, "rb").read()).hexdigest()}
json.dump(manifest, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
print("imported", len(files), "fixtures")
