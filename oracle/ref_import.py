"""Import the UNMODIFIED reference (FocoosAI/focoos @ /root/reference) on CPU, in THIS container only.

TEST INFRASTRUCTURE — not product code.  Used by `oracle/gen_golden.py` to produce the committed
fixtures under `tests/golden/` and by `tests/test_oracle_vs_reference.py` (skipped when
/root/reference is absent, e.g. on the GPU box).  Recipe = SURVEY.md Appendix C:
  * put /root/reference on sys.path (read-only tree → no bytecode),
  * stub the absent pure-python deps with permissive dummy modules,
  * patch importlib.metadata.version("focoos").
Nothing in the tensor path touches a stub.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.metadata
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FOCOOS_REFERENCE_ROOT", "/root/reference")

_STUB_TOPLEVEL = (
    "pycocotools", "supervision", "fvcore", "termcolor", "colorama", "orjson", "dotenv",
    "onnxruntime", "matplotlib", "IPython", "faster_coco_eval", "gradio", "shapely", "onnx",
    "tensorrt", "cv2", "tensorboardX", "onnxslim", "onnxscript",
)


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy_class(name)

    def __call__(cls, *a, **k):
        try:
            return super().__call__(*a, **k)
        except TypeError:
            return _make_dummy_class("inst")


def _make_dummy_class(name):
    return _DummyMeta(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: self})


class _DummyModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            v = _make_dummy_class(name)
        elif name in ("colored", "cprint", "init", "load_dotenv", "dumps", "loads"):
            v = lambda *a, **k: (a[0] if a else None)  # noqa: E731
        else:
            v = _DummyModule(self.__name__ + "." + name)
            sys.modules[v.__name__] = v
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _DummyModule(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "focoos"))


def install():
    """Idempotently make `import focoos` resolve to the reference tree."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    missing = []
    for n in _STUB_TOPLEVEL:
        try:
            __import__(n)
        except Exception:
            missing.append(n)
    sys.meta_path.append(_StubFinder(missing))
    _orig_version = importlib.metadata.version

    def _version(name):
        if name == "focoos":
            return "0.25.0"
        return _orig_version(name)

    importlib.metadata.version = _version
    if REFERENCE_ROOT not in sys.path:
        # appended, not prepended: the reference tree has its own top-level `tests` package, and multiprocessing's spawn hands this
        # process's sys.path to its children - prepending made `tests.<worker module>` resolve to the reference's package there
        sys.path.append(REFERENCE_ROOT)
    _installed = True


def get_reference_model(name: str, **kwargs):
    """ModelManager.get(name) with weights_uri=None (no network); returns the reference FocoosModel."""
    install()
    import torch
    from focoos.model_manager import ModelManager
    from focoos.model_registry.model_registry import ModelRegistry

    mi = ModelRegistry.get_model_info(name)
    mi.weights_uri = None
    torch.manual_seed(0)
    fm = ModelManager.get(name, model_info=mi, **kwargs)
    fm.model.eval()
    return fm
