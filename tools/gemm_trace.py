#!/usr/bin/env python
"""Per-CTA phase timeline of the CTA-pair GEMM (clock64 stamps). Dev tool: python tools/gemm_trace.py M N K epi"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import lib as L  # noqa: E402

lib = L.load()
M, N, K, epi = (int(x) for x in sys.argv[1:5])
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
trace = torch.zeros(148 * 16, device="cuda", dtype=torch.int64)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    L.check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, epi, 1, out.data_ptr(),
                               out.data_ptr() if epi == 2 else None, None, 0, 0, 0, st))


for _ in range(3):
    run()
lib.mmcm_debug_set_gemm_trace(C.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
lib.mmcm_debug_set_gemm_trace(None)
t = trace.view(148, 16).cpu()
names = ["entry", "prologue", "first data", "tile0 mma issued", "tile0 acc ready", "tile0 epi done", "last mma issued",
         "last acc ready", "last epi done", "exit"]
t0 = t[:, 0].min().item()
print(f"M={M} N={N} K={K} epi={epi}: clock64 cycles relative to each CTA's own entry (leader CTAs 0, 2, 72, 146; peer 1)")
for cta in (0, 2, 72, 146, 1):
    row = t[cta]
    print(f"cta {cta:3d}: " + "  ".join(f"{names[i]}={row[i].item() - row[0].item() if row[i].item() else -1}" for i in range(10)))
for cta in (0, 2, 72):
    row = t[cta]
    print(f"cta {cta}: epilogue warp 0, tile 0: ldtm issue->ready {row[11].item() - row[10].item()}  chunk0 math+store "
          f"{row[12].item() - row[11].item()}  chunk1 total {row[13].item() - row[12].item()}  "
          f"(acc ready -> first ldtm {row[10].item() - row[4].item()})  chunk0: 2xLDTM+pack {row[14].item() - row[10].item()} "
          f"store passes {row[12].item() - row[14].item()}")
ex = t[:, 9] - t0
print("exit stamp over CTAs: min", ex.min().item(), "max", ex.max().item(), " entry spread", (t[:, 0] - t0).max().item())
