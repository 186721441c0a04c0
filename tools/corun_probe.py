#!/usr/bin/env python
"""Do a GEMM stream and a LayerNorm / attention stream actually overlap on the SMs?  times alone vs together."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import lib as L  # noqa: E402

lib = L.load()
M, N, K = 51200, 3072, 768
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
rows, D = 78848, 512
x = torch.randn(rows, D, device="cuda")
g = torch.ones(D, device="cuda")
b = torch.zeros(D, device="cuda")
h = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
T, H = 77, 8
qkv = torch.randn(1024 * T, 3 * H * 64, device="cuda").bfloat16()
att = torch.empty(1024 * T, H * 64, device="cuda", dtype=torch.bfloat16)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def gemms(n, st):
    for _ in range(n):
        L.check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 1, 1, out.data_ptr(), None, None, 0, 0,
                                   0, C.c_void_p(st.cuda_stream)))


def lns(n, st):
    for _ in range(n):
        L.check(lib.mmcm_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5, rows, D, h.data_ptr(), None,
                                   C.c_void_p(st.cuda_stream)))


def atts(n, st):
    for _ in range(n):
        L.check(lib.mmcm_attention(qkv.data_ptr(), None, 1024, T, H, 1, att.data_ptr(), C.c_void_p(st.cuda_stream)))


def timed(fa, fb):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_stream(torch.cuda.current_stream())
    sb.wait_stream(torch.cuda.current_stream())
    if fa:
        fa(sa)
    if fb:
        fb(sb)
    torch.cuda.current_stream().wait_stream(sa)
    torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for name, other, n_other in (("layernorm", lns, 200), ("attention", atts, 60)):
    timed(lambda s: gemms(3, s), lambda s: other(3, s))
    tg = timed(lambda s: gemms(30, s), None)
    to = timed(None, lambda s: other(n_other, s))
    tb = timed(lambda s: gemms(30, s), lambda s: other(n_other, s))
    print(f"gemm alone {tg:.2f} ms | {name} alone {to:.2f} ms | together {tb:.2f} ms (sum {tg + to:.2f}, max {max(tg, to):.2f})")
