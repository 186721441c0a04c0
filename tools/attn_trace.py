#!/usr/bin/env python
"""Phase stamps of the tcgen05 attention kernel (2nd tile of CTA 0) + timing of both attention kernels. Dev tool."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
from mmcm_b200 import arch as A, lib as L  # noqa: E402

lib = L.load()
eng = P.Engine(A.CLIP_B32, A.HEAD_FUSION, 5, 512, 0, 0)   # only to reach set_option (process-wide switches)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for B, T, H, causal in ((1024, 77, 8, 1), (1024, 50, 12, 0)):
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").bfloat16()
    out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    for impl in (0, 1):
        eng.set_option("attention_impl", impl)
        for _ in range(3):
            L.check(lib.mmcm_attention(qkv.data_ptr(), None, B, T, H, causal, out.data_ptr(), st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            L.check(lib.mmcm_attention(qkv.data_ptr(), None, B, T, H, causal, out.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B} T={T} H={H} causal={causal} impl={impl}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
    eng.set_option("attention_impl", 0)
    trace = torch.zeros(16, device="cuda", dtype=torch.int64)
    lib.mmcm_debug_set_gemm_trace(C.c_void_p(trace.data_ptr()))
    L.check(lib.mmcm_attention(qkv.data_ptr(), None, B, T, H, causal, out.data_ptr(), st))
    torch.cuda.synchronize()
    lib.mmcm_debug_set_gemm_trace(None)
    t = trace.cpu().tolist()
    print("  tile stamps (cycles): wait S", t[1] - t[0], "| pass1", t[2] - t[1], "| pass2", t[3] - t[2], "| wait O", t[4] - t[3],
          "| epilogue", t[5] - t[4], "| total", t[5] - t[0])
