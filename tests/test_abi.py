"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/sylph_b200.h declares.
(No compute calls here: they need a device and live in the `-m gpu` tests.)"""
import ctypes
import os
import re

from sylph_few_shot_detection_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(REPO, "include", "sylph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sylph_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared_functions()
    assert len(declared) >= 18
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_library_loads_and_exports_every_declared_symbol():
    _lib.build()
    lib = _lib.load()
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/sylph_b200.h but not exported"
    assert b"sm_100a" in lib.sylph_version()


def test_model_config_struct_layout_matches_header():
    text = open(os.path.join(REPO, "include", "sylph_b200.h")).read()
    body = text[text.index("typedef struct sylph_model_config {"):text.index("} sylph_model_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|float)\s+([a-z_0-9]+)(?:\[3\])?;", body)
    assert fields == [f[0] for f in _lib.ModelConfig._fields_]
    assert ctypes.sizeof(_lib.ModelConfig) == 4 * (8 + 3 + 6 + 7 + 6 + 1)


def test_loss_config_struct_layout_matches_header():
    text = open(os.path.join(REPO, "include", "sylph_b200.h")).read()
    body = text[text.index("typedef struct sylph_loss_config {"):text.index("} sylph_loss_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|float)\s+([a-z_0-9]+)(?:\[4\])?;", body)
    assert fields == [f[0] for f in _lib.LossConfig._fields_]
    assert ctypes.sizeof(_lib.LossConfig) == 4 * (5 + 4)
    assert f"#define SYLPH_LOSS_SUMS {_lib.LOSS_SUMS}" in text and f"#define SYLPH_BACKGROUND_ID {_lib.BACKGROUND_ID}" in text


def test_codegen_tensors_struct_layout_matches_header():
    text = open(os.path.join(REPO, "include", "sylph_b200.h")).read()
    body = text[text.index("typedef struct sylph_codegen_tensors {"):text.index("} sylph_codegen_tensors;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\bfloat\*\s+([a-z_0-9]+)(?:\[SYLPH_CG_MAX_TOWER\])?;", body)
    assert fields == [f[0] for f in _lib.CodegenTensors._fields_]
    assert f"#define SYLPH_CG_MAX_TOWER {_lib.CG_MAX_TOWER}" in text
    assert ctypes.sizeof(_lib.CodegenTensors) == 8 * (4 * _lib.CG_MAX_TOWER + 8)


def test_tower_tensors_struct_layout_matches_header():
    text = open(os.path.join(REPO, "include", "sylph_b200.h")).read()
    body = text[text.index("typedef struct sylph_tower_tensors {"):text.index("} sylph_tower_tensors;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\bfloat\*\s+([a-z_0-9]+)\[SYLPH_CG_MAX_TOWER\];", body)
    assert fields == [f[0] for f in _lib.TowerTensors._fields_]
    assert ctypes.sizeof(_lib.TowerTensors) == 8 * 4 * _lib.CG_MAX_TOWER


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = ctypes.c_void_p()
    mc = _lib.ModelConfig()
    rc = lib.sylph_create(ctypes.byref(h), 0, ctypes.byref(mc))
    assert rc != 0 and not h, "there must be no CPU fallback"
    import pytest
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runtime import Engine
    with pytest.raises(RuntimeError):
        Engine(coco_meta_fcos_cfg(), 0)


def test_plain_c_client_compiles_against_the_header_and_runs(tmp_path):
    """include/sylph_b200.h is valid C99 (-pedantic -Werror) and the library links from a plain-C host
    (examples/c_client.c): without a device the client reports the failed sylph_create and exits 0."""
    import shutil
    import subprocess
    import torch
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    _lib.build()
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(REPO, "include"),
                    os.path.join(REPO, "examples", "c_client.c"), "-L", libdir, "-lsylph_b200", f"-Wl,-rpath,{libdir}", "-o", exe],
                   check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "sm_100a" in out and f"IPC handle: {_lib.IPC_HANDLE_BYTES} bytes" in out
    if not torch.cuda.is_available():
        assert "no CPU fallback" in out
    else:
        assert "weights not finalized" in out and "single-rank exchange: status 0, timed_out 0, rows 0" in out
