"""`src.models` of the reference (R/src/models/__init__.py:3-6), answered by the B200-native scoring path.

Exports the same two names with the same constructors, state-dict keys and
`forward(input_ids, attention_mask, pixel_values, text_present, image_present, labels=None) -> {"loss", "logits"}`
(R/src/models/fusion.py:83-95,157-165,229; R/src/models/multitask.py:40-52,156-164,227).  Reached either through the
import hook in shim/sitecustomize.py (PYTHONPATH=<repo>/shim, no edits to the reference) or by replacing the
reference's `src/models/` directory with this one.
"""
import importlib.util
import os
import sys

_REPO = os.environ.get("MMCM_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))))
_PKG_DIR = os.path.join(_REPO, "multimodal-content-moderation_b200")
_PKG_NAME = "mmcm_b200"   # the directory name has hyphens: registered under an importable name


def _load_package():
    if _PKG_NAME in sys.modules:
        return sys.modules[_PKG_NAME]
    init = os.path.join(_PKG_DIR, "__init__.py")
    if not os.path.exists(init):
        raise ImportError(f"B200 scoring package not found at {_PKG_DIR} (set MMCM_B200_ROOT to the repo root)")
    spec = importlib.util.spec_from_file_location(_PKG_NAME, init, submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load_package()
MultiModalFusionClassifier = _pkg.MultiModalFusionClassifier
MultiTaskClassifier = _pkg.MultiTaskClassifier
FocalWithLogitsLoss = _pkg.FocalWithLogitsLoss     # R/src/models/fusion.py:16-52 (importable from src.models.fusion)

__all__ = ["MultiModalFusionClassifier", "MultiTaskClassifier"]
