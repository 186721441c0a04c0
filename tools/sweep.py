#!/usr/bin/env python
"""Run bench.py over a list of flag sets and print one compact line each. Dev tool.
usage: python tools/sweep.py "--micro-batch 256 --l2-persist-mb 60" "--micro-batch 1024" ..."""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
base = [sys.executable, os.path.join(root, "bench.py"), "--steps", "12", "--warmup", "4", "--no-cpu-baseline", "--no-e2e"]
for flags in sys.argv[1:]:
    r = subprocess.run(base + flags.split(), capture_output=True, text=True, env=dict(os.environ, MMCM_DEBUG="1"))
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    note = [l for l in r.stderr.splitlines() if "persisting" in l]
    if line:
        d = json.loads(line[-1])
        print(f"{flags:45s} {d['value']:9.0f} samples/s  {d['ms_per_step']:7.2f} ms  gemm {d['roofline']['achieved']:6.0f} TF/s  {note[:1]}")
    else:
        print(f"{flags:45s} FAILED: {r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else 'no output'}")
