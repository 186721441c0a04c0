"""Numerics experiment (CPU, torch): would folding LayerNorm into the consumer GEMM keep parity?

    python tests/experiments/ln_fold_numerics.py

Three emulations of the CLIP towers on the golden cases, all compared with the fp32 oracle:
  bf16    : the shipped design -- h = bf16(LN(x)) is the GEMM operand, fp32 accumulate, fp32 residual stream
  fold    : LN(x) W^T = rstd * (bf16(x) @ bf16(W*gamma)^T - mu * s) + (W beta + b), s = rowsum(bf16(W*gamma)),
            mu / rstd from the fp32 residual (DESIGN.md section 8, item 1 alternative)
Not a test (no assertions): it prints logit errors so that the decision is recorded with numbers.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import build_case, oracle_forward  # noqa: E402
from oracle import scoring_oracle as orc  # noqa: E402

MODE = "fp32"


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def ln_linear(x, sd, ln, lins, eps):
    """[LN(x) @ W^T + b for each Linear in `lins`] under the current MODE."""
    g, b = sd[ln + ".weight"], sd[ln + ".bias"]
    outs = []
    if MODE == "fp32":
        h = F.layer_norm(x, (x.shape[-1],), g, b, eps)
        return [F.linear(h, sd[p + ".weight"], sd[p + ".bias"]) for p in lins]
    if MODE == "bf16":
        h = bf(F.layer_norm(x, (x.shape[-1],), g, b, eps))
        return [F.linear(h, bf(sd[p + ".weight"]), sd[p + ".bias"]) for p in lins]
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xa = bf(x)
    for p in lins:
        W, bias = sd[p + ".weight"], sd[p + ".bias"]
        Wg = bf(W * g[None, :])
        s = Wg.sum(dim=1)
        outs.append(rstd * (F.linear(xa, Wg) - mu * s[None, :]) + (W @ b + bias))
    return outs


def lin(x, sd, p):
    if MODE == "fp32":
        return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])
    return F.linear(bf(x), bf(sd[p + ".weight"]), sd[p + ".bias"])


def encoder(x, sd, prefix, layers, heads, eps, act, allow, stages=None, tag=""):
    q8 = (lambda t: t) if MODE == "fp32" else bf
    for i in range(layers):
        p = f"{prefix}encoder.layers.{i}."
        q, k, v = ln_linear(x, sd, p + "layer_norm1", [p + "self_attn.q_proj", p + "self_attn.k_proj",
                                                       p + "self_attn.v_proj"], eps)
        a = orc._attention(q8(q), q8(k), q8(v), heads, allow)
        x = x + lin(a, sd, p + "self_attn.out_proj")
        (h,) = ln_linear(x, sd, p + "layer_norm2", [p + "mlp.fc1"], eps)
        x = x + lin(q8(orc._act(h, act)), sd, p + "mlp.fc2")
    return x


def main():
    global MODE
    orc._encoder = encoder                      # the towers call _encoder through the module namespace
    for name in ("clip_fusion_default", "clip_fusion_hardened", "clip_mtl_h256_hardened"):
        kind, a, kw, sd, batch, gold = build_case(name)
        res = {}
        for MODE in ("fp32", "bf16", "fold"):
            with torch.no_grad():
                res[MODE] = oracle_forward(kind, a, sd, batch)
        ref = res["fp32"]
        print(f"{name:28s} logit std {ref.std():.4f} | max-abs error  bf16 design {(res['bf16'] - ref).abs().max():.2e}"
              f"   LN folded {(res['fold'] - ref).abs().max():.2e}")


if __name__ == "__main__":
    main()
