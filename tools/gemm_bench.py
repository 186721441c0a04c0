#!/usr/bin/env python
"""Times mmcm_gemm_bf16 on the tower shapes (CUDA events, inputs rotated through >L2 of buffers).  Dev tool."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import lib as L  # noqa: E402

lib = L.load()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
only = sys.argv[2] if len(sys.argv) > 2 else ""
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
SHAPES = []
for name, T, D, F in (("text", 77, 512, 2048), ("vision", 50, 768, 3072)):
    M = mb * T
    SHAPES += [(f"{name}.qkv", M, 3 * D, D, 0), (f"{name}.out", M, D, D, 2), (f"{name}.fc1", M, F, D, 1),
               (f"{name}.fc2", M, D, F, 2)]
SHAPES.append(("vision.patch", mb * 49, 768, 3072, 3))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tot_t = tot_f = 0.0
for name, M, N, K, epi in SHAPES:
    if only and only not in name:
        continue
    nbuf = int(os.environ.get("NBUF", "4"))
    A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    if epi in (0, 1):
        out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    else:
        out = [torch.zeros(M + M // 49 + 8, N, device="cuda") for _ in range(nbuf)]
    pos = torch.randn(50, N, device="cuda")

    def run(i):
        o = out[i % nbuf]
        L.check(lib.mmcm_gemm_bf16(A[i % nbuf].data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, epi, 1, o.data_ptr(),
                                   o.data_ptr() if epi == 2 else None, pos.data_ptr() if epi == 3 else None,
                                   49 if epi == 3 else 0, 50 if epi == 3 else 0, impl, st))
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * M * N * K
    tot_t += us
    tot_f += fl
    print(f"{name:14s} M={M:6d} N={N:5d} K={K:5d} epi={epi}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
print(f"sum: {tot_t:.1f} us  {tot_f / tot_t / 1e6:.1f} TFLOP/s (back-to-back launches, includes launch gaps)")
