"""ORACLE (test infrastructure only).  Import the reference's own hot-path modules, UNMODIFIED, from /root/reference
on top of the dependency stand-ins in oracle/shims.  Only usable in the build container (the GPU box has no
/root/reference); used by oracle/make_golden.py to emit tests/golden/*.pt and by container-only tests."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SYLPH_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sylph", "modeling"))


def load():
    """Returns a namespace with the reference classes needed on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)  # last: the reference has its own top-level `tests` package
    # `sylph/modeling/meta_arch/__init__.py:10-11` also imports the RCNN meta-archs (out of scope, need more of
    # detectron2); register an empty package in its place and load the one-stage file by path.
    if "sylph.modeling.meta_arch" not in sys.modules:
        importlib.import_module("sylph")
        pkg = types.ModuleType("sylph.modeling.meta_arch")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "sylph", "modeling", "meta_arch")]
        sys.modules["sylph.modeling.meta_arch"] = pkg
    import sylph.modeling.code_generator  # noqa: F401  registers CodeGenerator / CodeGeneratorHead
    import sylph.modeling.meta_fcos  # noqa: F401  registers MetaFCOS
    importlib.import_module("sylph.modeling.code_generator.roi_encoder")  # registers ROIEncoder (meta_fcos_runner.py:70-71)
    name = "sylph.modeling.meta_arch.meta_one_stage_detector"
    if name not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REFERENCE_ROOT, "sylph", "modeling", "meta_arch", "meta_one_stage_detector.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    ns = types.SimpleNamespace()
    ns.meta_arch = sys.modules[name]
    ns.MetaOneStageDetector = ns.meta_arch.MetaOneStageDetector
    ns.code_generator = importlib.import_module("sylph.modeling.code_generator.code_generator")
    ns.cg_utils = importlib.import_module("sylph.modeling.code_generator.utils")
    ns.roi_encoder = importlib.import_module("sylph.modeling.code_generator.roi_encoder")
    ns.fcos = importlib.import_module("sylph.modeling.meta_fcos.fcos")
    ns.fcos_outputs = importlib.import_module("sylph.modeling.meta_fcos.fcos_outputs")
    ns.head_utils = importlib.import_module("sylph.modeling.meta_fcos.head_utils")
    return ns


def build_reference_model(cfg):
    ns = load()
    model = ns.MetaOneStageDetector(cfg)
    model.eval()
    return model
