"""The oracle is pinned against outputs of the REAL reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference's own classes) and, piecewise, against Hugging Face's modules
(the third-party code the reference delegates its encoder arithmetic to).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, build_case, oracle_forward

from oracle import scoring_oracle as orc


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    kind, a, kw, sd, batch, gold = build_case(name)
    stages = {}
    with torch.no_grad():
        logits = oracle_forward(kind, a, sd, batch, stages).numpy()
    # fp32 vs fp32: differences are summation-order only
    np.testing.assert_allclose(logits, gold["logits"], rtol=2e-4, atol=2e-4)
    for key in ("text_pooled", "vision_pooled", "text_feat", "vision_feat"):
        if key in gold:
            np.testing.assert_allclose(stages[key].numpy(), gold[key], rtol=2e-4, atol=2e-4, err_msg=key)
    loss = orc.bce_loss(torch.from_numpy(logits), torch.from_numpy(gold["labels"])).item()
    if kind == "fusion":
        assert abs(loss - float(gold["loss"])) < 1e-4


def test_edge_rows_behave_like_the_reference_measurements():
    """absent modalities make the logits independent of that modality's input (Fusion)."""
    kind, a, kw, sd, batch, gold = build_case("clip_fusion_hardened")
    with torch.no_grad():
        base = oracle_forward(kind, a, sd, batch)
        b2 = dict(batch)
        px = batch["pixel_values"].clone()
        px[1] = torch.randn_like(px[1])            # row 1: image absent -> pixels irrelevant
        ids = batch["input_ids"].clone()
        ids[0, 1:5] = 17                            # row 0: text absent -> ids irrelevant
        b2["pixel_values"], b2["input_ids"] = px, ids
        alt = oracle_forward(kind, a, sd, b2)
    assert torch.equal(base[1], alt[1])
    assert torch.equal(base[0], alt[0])
    assert torch.equal(base[3:], alt[3:])


def test_clip_eos_pooling_rules():
    """First EOS, else index 0; legacy eos_id==2 -> argmax(ids)  (HF clip :564-584)."""
    kind, a, kw, sd, batch, gold = build_case("clip_fusion_hardened")
    ids = batch["input_ids"][:4].clone()
    S = ids.shape[1]
    ids[0] = torch.arange(1, S + 1)                # no EOS -> row 0
    ids[1, :] = 5; ids[1, 9] = a.eos_id            # EOS at 9
    ids[2, :] = 5; ids[2, 1] = a.eos_id; ids[2, 30] = a.eos_id   # first of two
    stages = {}
    with torch.no_grad():
        pooled = orc.clip_text_pooled(sd, "backbone.text_model.", ids, None, a.eos_id, stages=stages)
        h = stages["text_layer11"]
        h = torch.nn.functional.layer_norm(h, (512,), sd["backbone.text_model.final_layer_norm.weight"],
                                           sd["backbone.text_model.final_layer_norm.bias"], 1e-5)
        legacy = orc.clip_text_pooled(sd, "backbone.text_model.", ids, None, 2)
    assert torch.allclose(pooled[0], h[0, 0]) and torch.allclose(pooled[1], h[1, 9]) and torch.allclose(pooled[2], h[2, 1])
    assert torch.allclose(legacy[0], h[0, S - 1])  # argmax of 1..S is the last position


def test_sequence_too_long_raises_like_hf():
    kind, a, kw, sd, batch, gold = build_case("clip_fusion_hardened")
    ids = torch.zeros(1, 78, dtype=torch.long)
    with pytest.raises(ValueError, match="Sequence length must be less than max_position_embeddings"):
        orc.clip_text_pooled(sd, "backbone.text_model.", ids, None)


def test_oracle_towers_match_huggingface_modules():
    """Direct check against transformers' CLIPModel (present in this image): same state dict, same inputs."""
    transformers = pytest.importorskip("transformers")
    from mmcm_b200 import arch as A, synthetic as syn
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=5, hardened=True)
    hf = transformers.CLIPModel(transformers.CLIPConfig()).eval()
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    missing = hf.load_state_dict(bsd, strict=False)
    assert not missing.unexpected_keys
    assert all("position_ids" in k for k in missing.missing_keys)
    batch = syn.make_inputs(a, 8, seed=3, edge_rows=True)
    with torch.no_grad():
        t_ref = hf.text_model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).pooler_output
        v_ref = hf.vision_model(pixel_values=batch["pixel_values"]).pooler_output
        t = orc.clip_text_pooled(sd, "backbone.text_model.", batch["input_ids"], batch["attention_mask"], a.eos_id)
        v = orc.clip_vision_pooled(sd, "backbone.vision_model.", batch["pixel_values"], a.patch)
    assert (t - t_ref).abs().max().item() < 2e-4
    assert (v - v_ref).abs().max().item() < 2e-4
