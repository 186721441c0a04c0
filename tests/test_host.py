"""Host-side logic that needs no GPU: the C-ABI library loads and exports what include/mmcm.h declares, the
drop-in modules expose the reference's state-dict keys, and CPU use fails loudly."""
import pytest
import torch

from conftest import TASKS


def test_library_loads_and_exports_every_declared_symbol():
    from mmcm_b200 import lib as L
    syms = L.declared_symbols()
    assert {"mmcm_create", "mmcm_forward", "mmcm_forward_host", "mmcm_load_weight", "mmcm_finalize_weights",
            "mmcm_gemm_bf16", "mmcm_attention", "mmcm_layernorm"} <= set(syms)
    lib = L.load()
    for s in syms:
        assert hasattr(lib, s)
    assert b"sm_100a" in lib.mmcm_version()


def test_library_contains_blackwell_instructions():
    """cuobjdump evidence that the GEMM is tcgen05/TMA (SASS: UTCHMMA / UTMALDG / LDTM), not mma.sync."""
    import shutil
    import subprocess
    from mmcm_b200 import build as B, lib as L
    L.load()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_without_gpu_fails_loudly():
    import mmcm_b200 as P
    from mmcm_b200 import arch as A
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        P.Engine(A.CLIP_B32, A.HEAD_FUSION, 5)


def test_modules_mirror_reference_state_dict_keys():
    import mmcm_b200 as P
    m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
    keys = list(m.state_dict().keys())
    assert len(keys) == 416    # measured on the reference
    for k in ("backbone.text_model.encoder.layers.11.self_attn.q_proj.weight", "backbone.visual_projection.weight",
              "backbone.logit_scale", "proj_t.weight", "gate.weight", "cls.0.weight", "cls.4.bias", "ln_fused.bias"):
        assert k in keys
    assert m.state_dict()["gate.weight"].shape == (512, 1026)
    mt = P.MultiTaskClassifier("openai/clip-vit-base-patch32", TASKS, head_hidden_dim=256)
    k2 = mt.state_dict().keys()
    assert "tower_txt.text_model.final_layer_norm.weight" in k2 and "heads.4.3.bias" in k2
    assert "shared_head.1.weight" in k2
    mt0 = P.MultiTaskClassifier("openai/clip-vit-base-patch32", TASKS)
    assert mt0.state_dict()["heads.0.weight"].shape == (1, 512)
    with pytest.raises(AssertionError):   # the reference asserts for AutoModel backends (multitask.py:81-88)
        P.MultiTaskClassifier("google/siglip2-base-patch16-224", TASKS, backend="siglip")


def test_cpu_forward_raises_no_fallback():
    import mmcm_b200 as P
    from mmcm_b200 import arch as A, synthetic as syn
    m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5).eval()
    batch = syn.make_inputs(A.CLIP_B32, 2, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(**batch)


def test_algorithmic_flops_match_survey():
    from mmcm_b200 import arch as A
    f = A.algorithmic_flops_per_sample(A.CLIP_B32, A.HEAD_FUSION)
    assert abs(f["total"] / 1e9 - 14.783) < 0.01
    f = A.algorithmic_flops_per_sample(A.CLIP_B32, A.HEAD_MTL, head_hidden_dim=256)
    assert abs(f["total"] / 1e9 - 14.781) < 0.01
    f = A.algorithmic_flops_per_sample(A.SIGLIP2_B16, A.HEAD_FUSION)
    assert abs(f["total"] / 1e9 - 46.447) < 0.05


def test_synthetic_inputs_follow_survey_spec():
    from mmcm_b200 import arch as A, synthetic as syn
    b = syn.make_inputs(A.CLIP_B32, 64, seed=1234)
    ids, mask = b["input_ids"], b["attention_mask"]
    assert ids.shape == (64, 77) and (ids[:, 0] == 49406).all()
    lens = mask.sum(1)
    assert lens.min() >= 3 and lens.max() <= 77
    for r in range(64):
        assert (ids[r, lens[r] - 1:] == 49407).all() and (ids[r, 1:lens[r] - 1] < 49406).all()
    s = syn.make_inputs(A.SIGLIP2_B16, 16, seed=1)
    assert s["input_ids"].shape == (16, 64)


def test_focal_loss_formula():
    """Evaluation-time focal loss (fusion.py:39-52): ce * (1 - p_t)^gamma * alpha_t, hard and soft targets."""
    import torch.nn.functional as F
    from mmcm_b200.modules import FocalWithLogitsLoss
    g = torch.Generator().manual_seed(0)
    z = torch.randn(32, 5, generator=g) * 3
    alpha = torch.tensor([0.25, 0.5, 0.75, 0.3, 0.6])
    for t in ((torch.rand(32, 5, generator=g) < 0.3).float(), torch.rand(32, 5, generator=g)):
        p = torch.sigmoid(z)
        ce = F.binary_cross_entropy_with_logits(z, t, reduction="none")
        ref = ce * (1 - (p * t + (1 - p) * (1 - t))) ** 2.0 * (alpha * t + (1 - alpha) * (1 - t))
        assert torch.allclose(FocalWithLogitsLoss(alpha, 2.0, "none")(z, t), ref, atol=1e-6)
        assert torch.allclose(FocalWithLogitsLoss(alpha, 2.0, "mean")(z, t), ref.mean(), atol=1e-6)
        assert torch.allclose(FocalWithLogitsLoss(None, 2.0, "sum")(z, t),
                              (ce * (1 - (p * t + (1 - p) * (1 - t))) ** 2.0).sum(), atol=1e-4)


def test_resolve_arch_rejects_unknown_encoders_and_knows_siglip_v1():
    """an unknown `encoder_name` used to fall through silently to a CLIP / SigLIP2 shape."""
    import pytest
    from mmcm_b200 import arch as A
    assert A.resolve_arch("openai/clip-vit-base-patch32", "clip") is A.CLIP_B32
    assert A.resolve_arch("/data/ckpt/my-finetune-of-clip-vit-base-patch16", "clip") is A.CLIP_B16
    assert A.resolve_arch("google/siglip-base-patch16-224", "siglip").vocab == 32000       # SigLIP v1 vocabulary
    assert A.resolve_arch("google/siglip2-base-patch16-224", "siglip").vocab == 256000
    with pytest.raises(ValueError, match="unknown encoder_name"):
        A.resolve_arch("facebook/some-other-model", "clip")
    with pytest.raises(ValueError, match="architecture but backend"):
        A.resolve_arch("openai/clip-vit-base-patch32", "siglip")
    A.register_arch("my/clip-l14", A.CLIP_B16)
    assert A.resolve_arch("my/clip-l14", "clip") is A.CLIP_B16
    assert A.IMAGE_NORM[A.BACKEND_SIGLIP] == ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
