#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram (cuobjdump -sass), grouped by issue pipe.
Usage: tools/sass_histogram.py <cubin|so|exe> [kernel-name-regex]"""
import re, subprocess, sys
from collections import Counter
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "LEA")       # LEA measured to share the IMAD pipe? (see profiles/intpipe)
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat and not pat.search(name):
        continue
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
    c = Counter(ops)
    tot = sum(c.values())
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
    imad = sum(v for k, v in c.items() if k.startswith("IMAD")) - wide
    print(f"{name}: {tot} instr; IMAD.WIDE/HI {wide}, other IMAD {imad}, rest {tot - wide - imad}")
    print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
