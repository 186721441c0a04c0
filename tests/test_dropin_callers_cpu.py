"""The drop-in, driven by the REFERENCE'S OWN caller code (`model(**batch)["logits"]` under no_grad).

`/root/reference` exists only in the build container (it does not travel to the GPU box), and the build container has
no GPU, so these tests run on CPU with the one thing that needs a GPU -- the module's `_logits`, i.e. the C-ABI call
`mmcm_forward` -- answered by the CPU oracle.  Everything around it is real: the reference's unmodified
`scripts/evaluate.py:evaluate` (R/scripts/evaluate.py:163-183), its `collate_fn` (R/src/data/dataset.py:171-193) and
`scripts/inference.py:MultiModalClassifier.predict` (R/scripts/inference.py:182-237) import `src.models` through
shim/sitecustomize.py and call the B200 classes exactly as they call their own.  The CUDA kernels behind `_logits`
are covered against the same oracle by the `-m gpu` parity tests.
"""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, TASKS, build_case, oracle_forward

REF = os.environ.get("MMCM_REF_PATH", "/root/reference")
SHIM = os.path.join(ROOT, "shim")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "evaluate.py")),
                                reason="the reference repository is not present on this machine")


def test_pythonpath_shim_redirects_src_models_without_editing_the_reference():
    """`PYTHONPATH=shim python scripts/evaluate.py`: src.models -> B200 classes, the rest of `src` -> the reference."""
    code = (
        "import importlib.util, sys\n"
        f"spec = importlib.util.spec_from_file_location('ref_eval', r'{REF}/scripts/evaluate.py')\n"
        "m = importlib.util.module_from_spec(spec); sys.argv = ['evaluate.py']; spec.loader.exec_module(m)\n"
        "import src, src.data\n"
        "print(m.MultiModalFusionClassifier.__module__, m.MultiTaskClassifier.__module__, m.collate_fn.__module__,\n"
        "      src.__file__, src.data.__file__)\n"
        f"spec = importlib.util.spec_from_file_location('ref_inf', r'{REF}/scripts/inference.py')\n"
        "m2 = importlib.util.module_from_spec(spec); spec.loader.exec_module(m2)\n"
        "print(m2.MultiModalFusionClassifier.__module__)\n")
    env = dict(os.environ, PYTHONPATH=SHIM)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    l1, l2 = r.stdout.strip().splitlines()[-2:]
    fus, mtl, coll, src_file, data_file = l1.split()
    assert fus == "mmcm_b200.modules" and mtl == "mmcm_b200.modules" and l2.strip() == "mmcm_b200.modules"
    assert coll == "src.data.dataset" and src_file.startswith(REF) and data_file.startswith(REF)
    # without the shim the same import gives the reference's own classes
    r0 = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=600,
                        env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    assert r0.returncode == 0 and r0.stdout.strip().splitlines()[-2].split()[0] == "src.models.fusion"


@pytest.fixture(scope="module")
def ref_callers():
    """The reference's scripts/evaluate.py and scripts/inference.py, imported with the shim hook installed."""
    sys.path.insert(0, SHIM)
    spec = importlib.util.spec_from_file_location("mmcm_shim_sitecustomize", os.path.join(SHIM, "sitecustomize.py"))
    hook = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hook)                       # installs the meta-path finder (idempotent)
    saved_argv, saved_path = sys.argv, list(sys.path)
    sys.argv = ["evaluate.py"]
    mods = {}
    try:
        for nm in ("evaluate", "inference"):
            sp = importlib.util.spec_from_file_location(f"mmcm_ref_{nm}", os.path.join(REF, "scripts", f"{nm}.py"))
            mods[nm] = importlib.util.module_from_spec(sp)
            sp.loader.exec_module(mods[nm])
        from src.data import collate_fn                 # the reference's own collate (R/src/data/dataset.py:171-193)
        mods["collate_fn"] = collate_fn
        yield mods
    finally:
        sys.argv = saved_argv
        sys.path[:] = saved_path
        sys.meta_path[:] = [f for f in sys.meta_path if type(f).__name__ != "_B200ModelsFinder"]
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


def _oracle_backed(module_cls, kind, a, kw, sd):
    """A B200 drop-in module on CPU whose C-ABI forward is answered by the oracle (see the module docstring)."""
    name = "openai/clip-vit-base-patch32"
    m = module_cls(name, num_labels=5, **kw) if kind == "fusion" else module_cls(name, TASKS, **kw)
    m.load_state_dict(sd, strict=True)
    calls = []

    def _logits(input_ids, attention_mask, pixel_values, text_present, image_present):
        calls.append(tuple(input_ids.shape))
        batch = dict(input_ids=input_ids, attention_mask=attention_mask, pixel_values=pixel_values,
                     text_present=text_present, image_present=image_present)
        return oracle_forward(kind, a, {k: v.detach() for k, v in m.state_dict().items()}, batch)
    m._logits = _logits
    return m.eval(), calls


@pytest.mark.parametrize("case", ["clip_fusion_hardened", "clip_mtl_h256_hardened"])
def test_reference_evaluate_loop_runs_on_the_dropin(ref_callers, case):
    """R/scripts/evaluate.py:evaluate(model, dataloader, device), unmodified, over a DataLoader built with the
    reference's collate_fn from per-sample dicts shaped like SocialHarmDataset.__getitem__ (dataset.py:166-175)."""
    from torch.utils.data import DataLoader
    from mmcm_b200 import synthetic as syn
    ev = ref_callers["evaluate"]
    kind, a, kw, sd, _, _ = build_case(case)
    cls = ev.MultiModalFusionClassifier if kind == "fusion" else ev.MultiTaskClassifier
    assert cls.__module__ == "mmcm_b200.modules"
    model, calls = _oracle_backed(cls, kind, a, kw, sd)
    N = 11
    full = syn.make_inputs(a, N, seed=321, edge_rows=True)
    labels = (torch.arange(N * 5).reshape(N, 5) % 3 == 0).float()
    items = [{"input_ids": full["input_ids"][i], "attention_mask": full["attention_mask"][i],
              "pixel_values": full["pixel_values"][i], "labels": labels[i],
              "text_present": full["text_present"][i].clone(), "image_present": full["image_present"][i].clone()}
             for i in range(N)]
    loader = DataLoader(items, batch_size=4, shuffle=False, collate_fn=ref_callers["collate_fn"])
    logits, got_labels = ev.evaluate(model, loader, "cpu")
    assert calls == [(4, 77), (4, 77), (3, 77)]                    # one forward per collated batch, ragged tail
    assert logits.dtype == np.float32 and logits.shape == (N, 5)
    np.testing.assert_array_equal(got_labels, labels.numpy())
    with torch.no_grad():
        want = oracle_forward(kind, a, sd, full).numpy()
    np.testing.assert_allclose(logits, want, rtol=1e-5, atol=1e-5)
    # `labels` travels in the collated batch (evaluate.py:174): the drop-in accepts it and reports the reference's loss
    out = model(**ref_callers["collate_fn"](items[:4]))
    assert set(out) == {"loss", "logits"} and out["loss"] is not None and torch.isfinite(out["loss"])


def test_reference_predict_runs_on_the_dropin(ref_callers, tmp_path):
    """R/scripts/inference.py:MultiModalClassifier.predict with the B200 module injected: B=1 batches, thresholds,
    `any_harmful` -- the record format of inference.py:218-232."""
    from PIL import Image
    inf = ref_callers["inference"]
    kind, a, kw, sd, _, _ = build_case("clip_fusion_hardened")
    model, calls = _oracle_backed(inf.MultiModalFusionClassifier, kind, a, kw, sd)

    class _Tok:                                   # the CLIP tokenizer's contract: pad to max_length with EOS
        def __call__(self, text, padding, truncation, max_length, return_tensors):
            ids = [49406] + [1000 + (ord(c) % 500) for c in text][: max_length - 2] + [49407]
            n = len(ids)
            ids = ids + [49407] * (max_length - n)
            mask = [1] * n + [0] * (max_length - n)
            return {"input_ids": torch.tensor([ids]), "attention_mask": torch.tensor([mask])}

    class _Proc:
        size = {"shortest_edge": 224}
        image_mean = [0.48145466, 0.4578275, 0.40821073]
        image_std = [0.26862954, 0.26130258, 0.27577711]

    clf = inf.MultiModalClassifier.__new__(inf.MultiModalClassifier)      # skip _load_model (hub downloads)
    clf.device, clf.model, clf.tokenizer, clf.img_processor = "cpu", model, _Tok(), _Proc()
    clf.class_names = ["racist", "sexist", "homophobe", "religion", "otherhate"]
    clf.thresholds = [0.5, 0.35, 0.6, 0.5, 0.45]
    clf.img_size = clf._get_img_size()
    img = tmp_path / "x.png"
    Image.fromarray((np.random.RandomState(0).rand(300, 260, 3) * 255).astype(np.uint8)).save(img)
    rec = clf.predict(text="some text", image_path=str(img), return_probs=True)
    assert calls[-1] == (1, 77)
    assert set(rec) == {"predictions", "any_harmful", "probabilities"} and set(rec["predictions"]) == set(clf.class_names)
    probs = np.array(rec["probabilities"])
    for i, nm in enumerate(clf.class_names):
        p = rec["predictions"][nm]
        assert p["label"] == bool(probs[i] >= clf.thresholds[i]) and abs(p["probability"] - probs[i]) < 1e-7
    assert rec["any_harmful"] == any(v["label"] for v in rec["predictions"].values())
    # text-only and image-only requests (presence flags) go through the same call
    rec_t = clf.predict(text="only text")
    rec_i = clf.predict(image_path=str(img))
    assert set(rec_t) == {"predictions", "any_harmful"} and set(rec_i) == {"predictions", "any_harmful"}
    # the reference's batched helper is a loop over predict (inference.py:256-270)
    outs = clf.predict_batch(["a", "b"], [str(img), None], batch_size=32)
    assert len(outs) == 2 and calls[-2:] == [(1, 77), (1, 77)]
