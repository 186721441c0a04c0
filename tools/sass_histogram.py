#!/usr/bin/env python
"""SASS opcode histogram of libmmcm.so per kernel: the tcgen05 / TMA / tensor-core mnemonics that prove which hardware
path each kernel uses (B200_PROFILING.md).  Runs in the build container (cuobjdump, no GPU).

    python tools/sass_histogram.py [libmmcm.so] > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal-content-moderation_b200", "libmmcm.so")
WANT = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCATOMSWS",
        "HMMA", "LDSM", "LDGSTS", "SYNCS", "MUFU.TANH", "MUFU.EX2", "UCGABAR", "ACQBULK", "FFMA", "STS", "LDS", "STG", "LDG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void mmcm::", "").replace("mmcm::", "")
        hist[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                hist[cur][w] += 1
                break
print(f"SASS opcode counts per kernel of {os.path.basename(lib)} (cuobjdump -sass; static instruction counts)")
print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tile load /"
      " store / reduce-add, HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, MUFU.TANH = tanh.approx\n")
for k, c in hist.items():
    if c["_total"] == 0:
        continue
    items = " ".join(f"{w}={c[w]}" for w in WANT if c[w])
    print(f"{k:64s} total={c['_total']:6d}  {items}")
