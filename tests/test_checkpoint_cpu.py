"""SURVEY 8f rank 3 on the CPU: checkpoint directory -> the right module (the reference's evaluate.py:89-160 flow with
its documented config gaps fixed forward), and the packed-weight-file header checks that need no GPU."""
import json
import os

import pytest
import torch


def _write_checkpoint(tmp_path, sd, config, fmt="safetensors"):
    run = tmp_path / "run"
    ck = run / "checkpoint-7"
    ck.mkdir(parents=True)
    (run / "inference_config.json").write_text(json.dumps(config))
    if fmt == "safetensors":
        from safetensors.torch import save_file
        save_file({k: v.contiguous() for k, v in sd.items()}, str(ck / "model.safetensors"))
    else:
        torch.save(sd, str(ck / "pytorch_model.bin"))
    return ck


BASE = {"encoder_name": "openai/clip-vit-base-patch32", "backend": "clip", "fusion_dim": 512,
        "class_names": ["racist", "sexist", "homophobe", "religion", "otherhate"],
        "thresholds": [0.35, 0.7, 0.75, 0.3, 0.6]}


@pytest.mark.parametrize("hh,fmt", [(256, "safetensors"), (0, "bin")])
def test_mtl_checkpoint_without_head_key_loads_as_mtl(tmp_path, hh, fmt):
    """train.py never writes "head" (SURVEY 8b i) and the callers never pass head_hidden_dim (ii): the reference
    rebuilds a Fusion model and load_state_dict fails; here the keys decide."""
    from mmcm_b200 import arch as A, checkpoint as ck, synthetic as syn
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.mtl_spec(a, 5, 512, hh), a, seed=3)
    sd["pos_weight"] = torch.arange(1, 6).float()
    sd["log_vars"] = torch.full((5,), 0.25)
    d = _write_checkpoint(tmp_path, sd, BASE, fmt)
    model, config = ck.load_checkpoint(d, device="cpu")
    assert type(model).__name__ == "MultiTaskClassifier" and config["head"] == "mtl"
    assert model._head_hidden_dim == hh and model.task_names == BASE["class_names"]
    assert torch.equal(model.pos_weight, sd["pos_weight"]) and torch.equal(model.log_vars.data, sd["log_vars"])
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert not model.training


def test_fusion_checkpoint_with_focal_buffers_loads(tmp_path):
    from mmcm_b200 import arch as A, checkpoint as ck, synthetic as syn
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=4)
    sd["criterion.alpha"] = torch.linspace(0.1, 0.9, 5)
    d = _write_checkpoint(tmp_path, sd, dict(BASE, head="fusion"))
    model, config = ck.load_checkpoint(d, device="cpu")
    assert type(model).__name__ == "MultiModalFusionClassifier" and model.loss_type == "focal"
    assert torch.equal(model.criterion.alpha, sd["criterion.alpha"])
    assert config["thresholds"] == BASE["thresholds"]


def test_checkpoint_errors_mirror_the_reference(tmp_path):
    from mmcm_b200 import arch as A, checkpoint as ck, synthetic as syn
    with pytest.raises(FileNotFoundError, match="inference_config.json"):
        ck.load_checkpoint(tmp_path / "nowhere" / "checkpoint-1", device="cpu")
    run = tmp_path / "run"
    (run / "checkpoint-1").mkdir(parents=True)
    (run / "inference_config.json").write_text(json.dumps(BASE))
    with pytest.raises(FileNotFoundError, match="model weights"):
        ck.load_checkpoint(run / "checkpoint-1", device="cpu")
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=4)
    with pytest.raises(ValueError, match="head="):
        ck.infer_model_spec(sd, dict(BASE, head="mtl"))
    with pytest.raises(ValueError, match="class names"):
        ck.infer_model_spec(sd, dict(BASE, class_names=["harmful"]))
    with pytest.raises(ValueError, match="neither"):
        ck.infer_model_spec({"foo": torch.zeros(1)}, BASE)


def test_packed_file_header_is_validated_without_a_gpu(tmp_path):
    from mmcm_b200 import lib as L
    import ctypes as C
    lib = L.load()
    cfg = L.MmcmConfig()
    bad = tmp_path / "bad.bin"
    bad.write_bytes(os.urandom(8192))
    with pytest.raises(ValueError, match="not a packed weight file"):
        L.check(lib.mmcm_packed_config(str(bad).encode(), C.byref(cfg)))
    with pytest.raises(ValueError, match="cannot open"):
        L.check(lib.mmcm_packed_config(str(tmp_path / "missing.bin").encode(), C.byref(cfg)))
    short = tmp_path / "short.bin"
    short.write_bytes(b"MMCMPK01")
    with pytest.raises(ValueError, match="too short"):
        L.check(lib.mmcm_packed_config(str(short).encode(), C.byref(cfg)))
