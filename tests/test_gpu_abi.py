"""C-ABI behaviour that the Python wrappers normally hide: status codes, error strings, weight bookkeeping."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(head_hidden=0):
    import mmcm_b200 as P
    from mmcm_b200 import arch as A
    return P.Engine(A.CLIP_B32, A.HEAD_FUSION, 5, 512, head_hidden, 0)


def test_weight_loading_errors():
    from mmcm_b200 import arch as A, lib as L, synthetic as syn
    eng = _engine()
    lib = eng.lib
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0)
    w = sd["proj_t.weight"].contiguous()
    # unknown key -> EINVAL (ValueError), wrong size -> EINVAL, finalize with missing tensors -> ESTATE
    assert lib.mmcm_load_weight(eng._h, b"not.a.key", w.data_ptr(), w.numel()) == L.EINVAL
    assert "unexpected key" in L.last_error()
    assert lib.mmcm_load_weight(eng._h, b"proj_t.weight", w.data_ptr(), w.numel() - 1) == L.EINVAL
    assert "size mismatch" in L.last_error()
    assert lib.mmcm_load_weight(eng._h, b"proj_t.weight", w.data_ptr(), w.numel()) == L.OK
    assert lib.mmcm_finalize_weights(eng._h) == L.ESTATE
    assert "missing" in L.last_error()
    # keys of the reference state dict that the path does not read are accepted and ignored
    for k in (b"backbone.logit_scale", b"pos_weight", b"backbone.text_model.embeddings.position_ids"):
        assert lib.mmcm_load_weight(eng._h, k, w.data_ptr(), 1) == L.OK
    # forward before finalize -> ESTATE
    ids = torch.zeros(2, 77, dtype=torch.long, device="cuda")
    px = torch.zeros(2, 3, 224, 224, device="cuda")
    f = torch.ones(2, device="cuda")
    out = torch.empty(2, 5, device="cuda")
    rc = lib.mmcm_forward(eng._h, ids.data_ptr(), None, px.data_ptr(), f.data_ptr(), f.data_ptr(), 2, 77, out.data_ptr(),
                          None, None)
    assert rc == L.ESTATE
    # device-resident source tensors are accepted as well as host ones
    eng.load_state_dict({k: v.cuda() for k, v in sd.items()})
    rc = lib.mmcm_forward(eng._h, ids.data_ptr(), None, px.data_ptr(), f.data_ptr(), f.data_ptr(), 2, 77, out.data_ptr(),
                          None, None)
    assert rc == L.OK
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    # attention_mask = NULL means all ones
    ones = torch.ones(2, 77, dtype=torch.long, device="cuda")
    out2 = torch.empty_like(out)
    assert lib.mmcm_forward(eng._h, ids.data_ptr(), ones.data_ptr(), px.data_ptr(), f.data_ptr(), f.data_ptr(), 2, 77,
                            out2.data_ptr(), None, None) == L.OK
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    eng.close()


def test_options_and_stages():
    from mmcm_b200 import arch as A, lib as L, synthetic as syn
    eng = _engine()
    a = A.CLIP_B32
    eng.load_state_dict(syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0))
    with pytest.raises(ValueError):
        eng.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        eng.set_option("gemm_impl", 7)
    b = {k: v.cuda() for k, v in syn.make_inputs(a, 8, seed=2).items()}
    eng.forward(b["input_ids"], b["attention_mask"], b["pixel_values"], b["text_present"], b["image_present"])
    assert eng.stage("text_pooled").numel() == 8 * 512 and eng.stage("vision_pooled").numel() == 8 * 768
    with pytest.raises(RuntimeError):     # needs debug_feats
        eng.stage("text_feat")
    with pytest.raises(ValueError):
        eng.stage("nonsense")
    n0 = eng.last_launch_count()
    eng.set_option("varlen_text", 0)
    eng.forward(b["input_ids"], b["attention_mask"], b["pixel_values"], b["text_present"], b["image_present"])
    assert eng.last_launch_count() == n0 - 1      # dense text: one embedding kernel instead of plan + packed embed
    eng.close()


def test_create_rejects_bad_configs():
    import mmcm_b200 as P
    from dataclasses import replace
    from mmcm_b200 import arch as A
    with pytest.raises(ValueError):
        P.Engine(replace(A.CLIP_B32, patch=30), A.HEAD_FUSION, 5)          # image % patch != 0
    with pytest.raises(ValueError):
        P.Engine(A.SIGLIP2_B16, A.HEAD_MTL, 5)                               # MTL + siglip: reference asserts too
    with pytest.raises(ValueError):
        P.Engine(A.CLIP_B32, A.HEAD_FUSION, 5, device=99)


def test_empty_batch_is_a_no_op():
    from mmcm_b200 import arch as A, synthetic as syn
    eng = _engine()
    a = A.CLIP_B32
    eng.load_state_dict(syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0))
    ids = torch.zeros(0, 77, dtype=torch.long, device="cuda")
    px = torch.zeros(0, 3, 224, 224, device="cuda")
    f = torch.zeros(0, device="cuda")
    y = eng.forward(ids, ids, px, f, f)
    assert y.shape == (0, 5)
    eng.close()


@pytest.mark.parametrize("name", ["clip_fusion_hardened", "clip_mtl_h256_hardened", "siglip_fusion_hardened"])
def test_packed_weight_file_round_trip(name, tmp_path):
    """mmcm_save_packed -> mmcm_packed_config / mmcm_load_packed in a fresh handle: bit-identical logits without the
    fp32 state dict; a config mismatch or a flipped byte is refused."""
    import mmcm_b200 as P
    from conftest import build_case
    from test_gpu_forward import _make_module
    from mmcm_b200 import synthetic as syn
    kind, a, kw, sd, _, _ = build_case(name)
    m = _make_module(kind, a, kw, sd)
    batch = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, 9, seed=61, edge_rows=True).items()}
    ref = m(**batch)["logits"].clone()
    path = tmp_path / "packed.bin"
    m.save_packed(str(path))
    bf16_bytes = sum(v.numel() for v in sd.values()) * 2
    assert bf16_bytes < path.stat().st_size < 2.6 * bf16_bytes          # bf16 matrices + fp32 embeddings / heads
    scorer = P.PackedScorer(path, "cuda:0")
    out = scorer(**batch)
    assert out["loss"] is None and torch.equal(out["logits"], ref)
    ea = scorer.engine.arch                                             # rebuilt from the header (eps went through fp32)
    assert scorer.engine.num_outputs == ref.shape[1]
    assert (ea.backend, ea.image, ea.patch, ea.max_pos, ea.vocab, ea.text.hidden, ea.vision.hidden, ea.text.layers) == \
        (a.backend, a.image, a.patch, a.max_pos, a.vocab, a.text.hidden, a.vision.hidden, a.text.layers)
    # a handle with another configuration refuses the file
    other = "clip_mtl_h256_hardened" if name != "clip_mtl_h256_hardened" else "clip_fusion_hardened"
    kind2, a2, kw2, sd2, _, _ = build_case(other)
    m2 = _make_module(kind2, a2, kw2, sd2)
    with pytest.raises(ValueError, match="different model configuration"):
        m2._ensure_engine(0).load_packed(str(path))
    # corruption is detected by the checksum
    raw = bytearray(path.read_bytes())
    raw[len(raw) // 2] ^= 0x40
    bad = tmp_path / "corrupt.bin"
    bad.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        P.PackedScorer(bad, "cuda:0")
