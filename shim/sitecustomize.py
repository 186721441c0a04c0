"""Zero-edit drop-in for the reference's callers.

    PYTHONPATH=/path/to/this/repo/shim python scripts/evaluate.py --checkpoint ... --test_csv ...
    PYTHONPATH=/path/to/this/repo/shim python scripts/inference.py --checkpoint ... --text ...

The reference's entry points put their own project root first on `sys.path`
(R/scripts/evaluate.py:17, R/scripts/inference.py:26, R/sagemaker/inference.py:72-73) and then run
`from src.models import MultiModalFusionClassifier, MultiTaskClassifier` (evaluate.py:26, inference.py:35).
Python imports a module named `sitecustomize` from PYTHONPATH at interpreter start-up; this one registers an import
hook that answers `src.models` (and its two sub-modules) with the B200 classes, while every other part of the
reference's `src` package (data, training, utils) still loads from the reference.  Nothing in the reference is edited.

The alternative without a hook: replace the reference's `src/models/` directory by `shim/src/models/`.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

_SHIM_DIR = os.path.dirname(os.path.abspath(__file__))
_TARGETS = {"src.models": True, "src.models.fusion": False, "src.models.multitask": False}


class _B200ModelsFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Answers `src.models*` from shim/src/models/__init__.py, whatever `src` package is on sys.path."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in _TARGETS or os.environ.get("MMCM_DROPIN", "1") == "0":
            return None
        is_pkg = _TARGETS[fullname]
        origin = os.path.join(_SHIM_DIR, "src", "models", "__init__.py")
        spec = importlib.machinery.ModuleSpec(fullname, self, origin=origin, is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.dirname(origin)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = os.path.join(_SHIM_DIR, "src", "models", "__init__.py")
        module.__file__ = origin
        with open(origin) as f:
            code = compile(f.read(), origin, "exec")
        exec(code, module.__dict__)


def install():
    if not any(isinstance(f, _B200ModelsFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _B200ModelsFinder())


install()
