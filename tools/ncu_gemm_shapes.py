#!/usr/bin/env python
"""Launch every encoder-GEMM shape of the CLIP-Fusion forward once inside an NVTX range, for
    ncu --set full --nvtx --nvtx-include "measure/" --clock-control none -o gpurun_out/r02_gemm_shapes python tools/ncu_gemm_shapes.py
Shapes are the ones the engine runs at batch 1024 (text chunks of 492 samples = 37 884 rows, vision chunk of 1024
samples = 51 200 rows) with the LN-fold epilogues.  Prints the launch order as JSON (tools/ncu_traffic.py joins it
with the report)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import lib as L  # noqa: E402

lib = L.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
order = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, fn, M, N, K, bytes_alg):
    for _ in range(2):
        fn()
    flush.zero_()                      # operands of the measured launch come from HBM, like inside the forward
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measure")
    fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    order.append({"name": name, "M": M, "N": N, "K": K, "flops": 2.0 * M * N * K, "algorithmic_bytes": bytes_alg})


for tower, rows, D, F in (("text", 492 * 77, 512, 2048), ("vision", 1024 * 50, 768, 3072)):
    M = rows
    x = torch.randn(M, D, device="cuda")
    xb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(D // 128, M, 2, device="cuda")
    L.check(lib.mmcm_prep_rows(x.data_ptr(), None, None, 1e-5, M, D, xb.data_ptr(), stats.data_ptr(), st))
    for N, nm, act in ((3 * D, "qkv", 0), (F, "fc1", 1)):
        W = (torch.randn(N, D, device="cuda") * D ** -0.5).bfloat16()
        b, cs = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        run(f"{tower}.{nm} (EPI_LNFOLD{'_ACT' if act else ''})",
            lambda: L.check(lib.mmcm_gemm_lnfold(xb.data_ptr(), W.data_ptr(), b.data_ptr(), stats.data_ptr(),
                                                 M, N, D, 1e-5, act, out.data_ptr(), st)),
            M, N, D, M * D * 2 + N * D * 2 + M * N * 2 + M * (D // 128) * 8)
        del out
    for K, nm in ((D, "out"), (F, "fc2")):
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(D, K, device="cuda") * K ** -0.5).bfloat16()
        b = torch.randn(D, device="cuda")
        run(f"{tower}.{nm} (EPI_RESID_STATS)",
            lambda: L.check(lib.mmcm_gemm_resid_stats(A.data_ptr(), W.data_ptr(), b.data_ptr(), M, D, K, x.data_ptr(),
                                                      xb.data_ptr(), stats.data_ptr(), st)),
            M, D, K, M * K * 2 + D * K * 2 + M * D * 10 + M * (D // 128) * 8)
        del A
    if tower == "vision":
        Mp, Kp = 1024 * 49, 3072
        A = torch.randn(Mp, Kp, device="cuda").bfloat16()
        W = (torch.randn(D, Kp, device="cuda") * 0.02).bfloat16()
        pos = torch.randn(50, D, device="cuda")
        xo = torch.zeros(1024 * 50, D, device="cuda")
        run("vision.patch (EPI_PATCH_F32)",
            lambda: L.check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), None, Mp, D, Kp, 3, 0, xo.data_ptr(), None,
                                               pos.data_ptr(), 49, 50, 0, st)),
            Mp, D, Kp, Mp * Kp * 2 + D * Kp * 2 + Mp * D * 4)
print("ORDER_JSON " + json.dumps(order))
