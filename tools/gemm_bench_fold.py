#!/usr/bin/env python
"""Per-shape timing of the LN-fold GEMMs against the kernels they replace (CUDA events, operands rotated through
buffers larger than L2).  Dev tool.

    python tools/gemm_bench_fold.py [samples_per_chunk=512] [iters=30]

For each tower (text 77 x 512 / 2048, vision 50 x 768 / 3072) prints, per layer step:
    old:  resid GEMM (L2 reduce-add) + LayerNorm pass          new:  mmcm_gemm_resid_stats
    old:  plain qkv / fc1 GEMM                                 new:  mmcm_gemm_lnfold
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import lib as L  # noqa: E402

lib = L.load()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
NB = 3


def timed(fn):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


tot_old = tot_new = 0.0
for name, T, D, F, act in (("text", 77, 512, 2048, 1), ("vision", 50, 768, 3072, 1)):
    M = mb * T
    x = [torch.randn(M, D, device="cuda") for _ in range(NB)]
    xb = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    stats = [torch.zeros(D // 128, M, 2, device="cuda") for _ in range(NB)]
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    for K, nm in ((D, "out"), (F, "fc2")):
        A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
        W = (torch.randn(D, K, device="cuda") * K ** -0.5).bfloat16()
        bias = torch.randn(D, device="cuda")
        t_res = timed(lambda i: L.check(lib.mmcm_gemm_bf16(A[i % NB].data_ptr(), W.data_ptr(), bias.data_ptr(), M, D, K, 2, 0,
                                                           x[i % NB].data_ptr(), x[i % NB].data_ptr(), None, 0, 0, 0, st)))
        t_ln = timed(lambda i: L.check(lib.mmcm_layernorm(x[i % NB].data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5, M, D,
                                                          xb[i % NB].data_ptr(), None, st)))
        t_new = timed(lambda i: L.check(lib.mmcm_gemm_resid_stats(A[i % NB].data_ptr(), W.data_ptr(), bias.data_ptr(), M, D, K,
                                                                  x[i % NB].data_ptr(), xb[i % NB].data_ptr(),
                                                                  stats[i % NB].data_ptr(), st)))
        fl = 2.0 * M * D * K
        print(f"{name}.{nm:4s} M={M:6d} N={D:5d} K={K:5d}  resid {t_res:7.1f} us + LN {t_ln:6.1f} us = {t_res + t_ln:7.1f}"
              f"   resid_stats {t_new:7.1f} us ({fl / t_new / 1e6:6.1f} TF/s, {M * D * 10 / t_new / 1e3:6.0f} GB/s of x traffic)")
        tot_old += t_res + t_ln
        tot_new += t_new
        del A
    for N, nm, a_ in ((3 * D, "qkv", 0), (F, "fc1", act)):
        W = (torch.randn(N, D, device="cuda") * D ** -0.5).bfloat16()
        bias, cs = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
        out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
        t_old = timed(lambda i: L.check(lib.mmcm_gemm_bf16(xb[i % NB].data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, D,
                                                           1 if a_ else 0, a_, out[i % NB].data_ptr(), None, None, 0, 0, 0, st)))
        t_new = timed(lambda i: L.check(lib.mmcm_gemm_lnfold(xb[i % NB].data_ptr(), W.data_ptr(), bias.data_ptr(),
                                                             stats[i % NB].data_ptr(), M, N, D, 1e-5, a_,
                                                             out[i % NB].data_ptr(), st)))
        fl = 2.0 * M * N * D
        print(f"{name}.{nm:4s} M={M:6d} N={N:5d} K={D:5d}  plain {t_old:7.1f} us ({fl / t_old / 1e6:6.1f} TF/s)"
              f"   lnfold {t_new:7.1f} us ({fl / t_new / 1e6:6.1f} TF/s)")
        tot_old += t_old
        tot_new += t_new
        del out
print(f"layer sums: old {tot_old:.1f} us  new {tot_new:.1f} us")
