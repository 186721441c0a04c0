"""Randomised mask / length / flag patterns on shallow towers (2 text + 2 vision layers of the real B/32 widths, so the
oracle runs in milliseconds): attention masks with holes and left padding, several EOS tokens, EOS-free rows, S from 1
to 77, arbitrary presence flags, legacy argmax pooling.  CUDA path vs oracle, both text modes, both attention kernels."""
from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _shallow(a):
    from mmcm_b200 import arch as A
    return replace(a, text=replace(a.text, layers=2), vision=replace(a.vision, layers=2))


def _random_batch(a, B, S, g, clip=True):
    ids = torch.randint(1, a.vocab - 2, (B, S), generator=g)
    mask = torch.ones(B, S, dtype=torch.long)
    for b in range(B):
        kind = int(torch.randint(0, 6, (1,), generator=g))
        L = int(torch.randint(1, S + 1, (1,), generator=g))
        if clip:
            if kind == 0:                                   # right padded, EOS then EOS-valued padding
                ids[b, L - 1:] = a.eos_id
                mask[b, L:] = 0
            elif kind == 1:                                 # holes in the mask
                ids[b, L - 1] = a.eos_id
                mask[b] = (torch.rand(S, generator=g) > 0.3).long()
            elif kind == 2:                                 # left padding
                ids[b, S - 1] = a.eos_id
                mask[b, : S - L] = 0
            elif kind == 3:                                 # several EOS, first one counts
                ids[b, L - 1] = a.eos_id
                ids[b, S - 1] = a.eos_id
            elif kind == 4:                                 # no EOS at all -> pooled row 0
                pass
            else:                                           # everything masked
                ids[b, L - 1] = a.eos_id
                mask[b] = 0
        else:
            mask[b, L:] = 0
            if kind == 1:
                mask[b] = (torch.rand(S, generator=g) > 0.3).long()
            if kind == 5:
                mask[b] = 0
    px = torch.randn(B, 3, a.image, a.image, generator=g)
    tp = (torch.rand(B, generator=g) > 0.25).float()
    ip = (torch.rand(B, generator=g) > 0.25).float()
    return {"input_ids": ids, "attention_mask": mask, "pixel_values": px, "text_present": tp, "image_present": ip}


@pytest.mark.parametrize("seed", range(6))
def test_random_masks_clip_fusion(seed):
    import mmcm_b200 as P
    from mmcm_b200 import arch as A, synthetic as syn
    from oracle import scoring_oracle as orc
    g = torch.Generator().manual_seed(1000 + seed)
    a = _shallow(A.CLIP_B32)
    if seed == 5:
        a = replace(a, eos_id=2)                            # legacy checkpoints: pooled row = argmax(ids)
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=seed, hardened=True)
    eng = P.Engine(a, A.HEAD_FUSION, 5, 512, 0, 0)
    eng.load_state_dict(sd)
    for S in (77, int(torch.randint(1, 77, (1,), generator=g)), 1):
        B = int(torch.randint(1, 40, (1,), generator=g))
        batch = _random_batch(a, B, S, g)
        with torch.no_grad():
            ref = orc.fusion_forward(sd, batch, "clip", a.patch, a.eos_id)
        d = {k: v.cuda() for k, v in batch.items()}
        outs = []
        for varlen, att in ((1, 1), (0, 2), (0, 1), (1, 0)):
            eng.set_option("varlen_text", varlen)
            eng.set_option("attention_impl", att)
            y = eng.forward(d["input_ids"], d["attention_mask"], d["pixel_values"], d["text_present"],
                            d["image_present"]).cpu()
            assert torch.isfinite(y).all()
            err = (y - ref).abs().max().item()
            assert err <= 0.05 * max(ref.std().item(), 0.5), f"seed {seed} S={S} B={B} varlen={varlen} att={att}: {err}"
            outs.append(y)
        assert torch.equal(outs[0], outs[2])                # packed == dense when both towers use the same attention kernel
    eng.set_option("attention_impl", 0)
    eng.close()


@pytest.mark.parametrize("seed", range(2))
def test_random_masks_siglip_fusion(seed):
    import mmcm_b200 as P
    from mmcm_b200 import arch as A, synthetic as syn
    from oracle import scoring_oracle as orc
    g = torch.Generator().manual_seed(2000 + seed)
    a = replace(_shallow(A.SIGLIP2_B16), vocab=4096)        # small vocabulary: the embedding table is irrelevant here
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=seed, hardened=True)
    eng = P.Engine(a, A.HEAD_FUSION, 5, 512, 0, 0)
    eng.load_state_dict(sd)
    for S in (64, int(torch.randint(1, 64, (1,), generator=g))):
        B = int(torch.randint(1, 12, (1,), generator=g))
        batch = _random_batch(a, B, S, g, clip=False)
        with torch.no_grad():
            ref = orc.fusion_forward(sd, batch, "siglip", a.patch, a.eos_id)
        d = {k: v.cuda() for k, v in batch.items()}
        for att in (0, 1):
            eng.set_option("attention_impl", att)
            y = eng.forward(d["input_ids"], d["attention_mask"], d["pixel_values"], d["text_present"],
                            d["image_present"]).cpu()
            err = (y - ref).abs().max().item()
            assert err <= 0.05 * max(ref.std().item(), 0.5), f"seed {seed} S={S} B={B} att={att}: {err}"
    eng.set_option("attention_impl", 0)
    eng.close()
