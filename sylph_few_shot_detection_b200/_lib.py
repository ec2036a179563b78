"""ctypes binding of libsylph_b200.so (C ABI declared in include/sylph_b200.h) and its in-tree build recipe.

The shared library is the product; this module only declares argument types.  There is no CPU fallback: if the
library is missing it is built with nvcc for sm_100a, and if that is impossible an ImportError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libsylph_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]

CODE_STRIDE = 257
DET_STRIDE = 9
IPC_HANDLE_BYTES = 64
NUM_LEVELS = 5
SLOT_SUPPORT, SLOT_QUERY = 0, 1


class ModelConfig(Structure):
    """Mirror of `sylph_model_config`."""
    _fields_ = [
        ("resnet_depth", c_int), ("num_cls_convs", c_int), ("num_box_convs", c_int), ("use_scale", c_int),
        ("thresh_with_ctr", c_int), ("box_quality", c_int), ("pre_nms_topk", c_int), ("post_nms_topk", c_int),
        ("inference_thresh", c_float), ("nms_thresh", c_float), ("prior_prob", c_float),
        ("pixel_mean", c_float * 3), ("pixel_std", c_float * 3),
        ("cg_tower_layers", c_int), ("cg_post_norm", c_int), ("cg_conv_l2_norm", c_int), ("cg_bias_layer", c_int),
        ("cg_bias_l2_norm", c_int), ("cg_use_bias", c_int), ("cg_has_conv_scale", c_int),
        ("generator", c_int), ("re_tok_convs", c_int), ("re_tok_fcs", c_int), ("re_layers", c_int),
        ("re_head_fcs", c_int), ("re_head_dim", c_int), ("cg_weight_layer", c_int),
    ]


class LossConfig(Structure):
    """Mirror of `sylph_loss_config` (training forward)."""
    _fields_ = [("focal_alpha", c_float), ("focal_gamma", c_float), ("center_sample", c_int), ("pos_radius", c_float),
                ("loc_loss_type", c_int), ("sizes_of_interest", c_int * 4)]


LOSS_SUMS = 5
BACKGROUND_ID = 100000
CG_MAX_TOWER = 4


class TowerTensors(Structure):
    """Mirror of `sylph_tower_tensors`: device fp32 pointers of the FCOS class tower's tensors (or their gradients)."""
    _fields_ = [("conv_w", c_void_p * CG_MAX_TOWER), ("conv_b", c_void_p * CG_MAX_TOWER),
                ("gn_w", c_void_p * CG_MAX_TOWER), ("gn_b", c_void_p * CG_MAX_TOWER)]


class CodegenTensors(Structure):
    """Mirror of `sylph_codegen_tensors`: device fp32 pointers of the code generator's trainable tensors (or their gradients)."""
    _fields_ = [("tower_w", c_void_p * CG_MAX_TOWER), ("tower_b", c_void_p * CG_MAX_TOWER),
                ("tower_gn_w", c_void_p * CG_MAX_TOWER), ("tower_gn_b", c_void_p * CG_MAX_TOWER),
                ("cls_w", c_void_p), ("cls_b", c_void_p), ("bias_w", c_void_p), ("bias_b", c_void_p),
                ("post_norm_w", c_void_p), ("post_norm_b", c_void_p), ("conv_scale", c_void_p), ("bias_scale", c_void_p)]


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(os.path.dirname(_PKG), "include", "sylph_b200.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources() if os.path.exists(s))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU; the .so stays in-tree (git-ignored, travels with gpurun)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH, os.path.join(_CSRC, "engine.cu")]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise ImportError(f"building libsylph_b200.so failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = ctypes.CDLL(LIB_PATH)
    if not hasattr(lib, "sylph_set_loss_box_branch"):   # a stale in-tree build from before the training backward
        build(force=True)
        lib = ctypes.CDLL(LIB_PATH)
    vp, ip, fp = c_void_p, POINTER(c_int), POINTER(c_float)
    lib.sylph_version.restype = c_char_p
    lib.sylph_version.argtypes = []
    lib.sylph_create.restype = c_int
    lib.sylph_create.argtypes = [POINTER(vp), c_int, POINTER(ModelConfig)]
    lib.sylph_destroy.restype = None
    lib.sylph_destroy.argtypes = [vp]
    lib.sylph_last_error.restype = c_char_p
    lib.sylph_last_error.argtypes = [vp]
    lib.sylph_set_precision.restype = c_int
    lib.sylph_set_precision.argtypes = [vp, c_int]
    lib.sylph_get_precision.restype = c_int
    lib.sylph_get_precision.argtypes = [vp]
    lib.sylph_load_tensor.restype = c_int
    lib.sylph_load_tensor.argtypes = [vp, c_char_p, vp, POINTER(c_int64), c_int]
    lib.sylph_finalize_weights.restype = c_int
    lib.sylph_finalize_weights.argtypes = [vp]
    lib.sylph_extract_features.restype = c_int
    lib.sylph_extract_features.argtypes = [vp, c_int, c_int, POINTER(vp), ip, ip, vp]
    lib.sylph_extract_features_u8.restype = c_int
    lib.sylph_extract_features_u8.argtypes = [vp, c_int, c_int, POINTER(vp), ip, ip, vp]
    lib.sylph_extract_features_normalized.restype = c_int
    lib.sylph_extract_features_normalized.argtypes = [vp, c_int, c_int, vp, c_int, c_int, vp]
    lib.sylph_extract_features_multi.restype = c_int
    lib.sylph_extract_features_multi.argtypes = [vp, c_int, ip, ip, POINTER(vp), c_int, ip, ip, vp]
    lib.sylph_import_features.restype = c_int
    lib.sylph_import_features.argtypes = [vp, c_int, c_int, c_int, c_int, POINTER(vp), ip, ip, vp]
    lib.sylph_set_image_sizes.restype = c_int
    lib.sylph_set_image_sizes.argtypes = [vp, c_int, c_int, ip, ip]
    lib.sylph_feature_shape.restype = c_int
    lib.sylph_feature_shape.argtypes = [vp, c_int, ip, ip, ip, ip, ip]
    lib.sylph_export_features.restype = c_int
    lib.sylph_export_features.argtypes = [vp, c_int, c_int, vp, vp]
    lib.sylph_generate_codes.restype = c_int
    lib.sylph_generate_codes.argtypes = [vp, c_int, c_int, fp, ip, c_int, ip, vp, vp, vp]
    lib.sylph_export_roi_features.restype = c_int
    lib.sylph_export_roi_features.argtypes = [vp, vp, vp]
    lib.sylph_normalize_codes.restype = c_int
    lib.sylph_normalize_codes.argtypes = [vp, vp, vp, c_int, vp]
    lib.sylph_exchange_create.restype = c_int
    lib.sylph_exchange_create.argtypes = [vp, c_int, c_int, c_int, vp]
    lib.sylph_exchange_connect.restype = c_int
    lib.sylph_exchange_connect.argtypes = [vp, vp]
    lib.sylph_normalize_codes_exchange.restype = c_int
    lib.sylph_normalize_codes_exchange.argtypes = [vp, vp, c_int, c_int, c_int, vp, vp]
    lib.sylph_exchange_poll.restype = c_int
    lib.sylph_exchange_poll.argtypes = [vp]
    lib.sylph_exchange_status.restype = c_int
    lib.sylph_exchange_status.argtypes = [vp, ip, POINTER(c_int64)]
    lib.sylph_exchange_destroy.restype = None
    lib.sylph_exchange_destroy.argtypes = [vp]
    lib.sylph_accumulate_codes.restype = c_int
    lib.sylph_accumulate_codes.argtypes = [vp, vp, c_int, ip, fp, vp, c_int, vp]
    lib.sylph_reduce_codes.restype = c_int
    lib.sylph_reduce_codes.argtypes = [vp, vp, c_int, c_int, fp, vp, vp]
    lib.sylph_detect.restype = c_int
    lib.sylph_detect.argtypes = [vp, c_int, vp, c_int, ip, vp, vp, c_int, vp]
    lib.sylph_detect_after.restype = c_int
    lib.sylph_detect_after.argtypes = [vp, c_int, vp, c_int, ip, vp, vp, c_int, vp, vp]
    lib.sylph_detect_poll.restype = c_int
    lib.sylph_detect_poll.argtypes = [vp]
    lib.sylph_export_head_output.restype = c_int
    lib.sylph_export_head_output.argtypes = [vp, c_int, c_int, vp, vp]
    lib.sylph_fcos_loss_sums.restype = c_int
    lib.sylph_fcos_loss_sums.argtypes = [vp, c_int, vp, c_int, POINTER(c_int64), POINTER(LossConfig), c_int, fp,
                                         POINTER(c_int64), ip, vp, vp, vp, vp, vp]
    lib.sylph_fcos_loss_finalize.restype = c_int
    lib.sylph_fcos_loss_finalize.argtypes = [vp, vp, vp, c_int, vp, vp]
    lib.sylph_fcos_cls_loss_backward.restype = c_int
    lib.sylph_fcos_cls_loss_backward.argtypes = [vp, c_int, c_int, POINTER(c_int64), POINTER(LossConfig), vp, vp, vp, c_int, vp, vp, vp]
    lib.sylph_codegen_backward.restype = c_int
    lib.sylph_codegen_backward.argtypes = [vp, c_int, c_int, ip, vp, vp, POINTER(CodegenTensors), POINTER(CodegenTensors), vp]
    lib.sylph_update_code_generator.restype = c_int
    lib.sylph_update_code_generator.argtypes = [vp]
    lib.sylph_update_code_generator_device.restype = c_int
    lib.sylph_update_code_generator_device.argtypes = [vp, POINTER(CodegenTensors), vp]
    lib.sylph_set_loss_box_branch.restype = c_int
    lib.sylph_set_loss_box_branch.argtypes = [vp, c_int]
    lib.sylph_set_training.restype = c_int
    lib.sylph_set_training.argtypes = [vp, c_int]
    lib.sylph_cls_tower_backward.restype = c_int
    lib.sylph_cls_tower_backward.argtypes = [vp, c_int, c_int, vp, POINTER(c_int64), POINTER(LossConfig), vp, vp, vp, c_int, vp,
                                             POINTER(TowerTensors), POINTER(TowerTensors), vp]
    lib.sylph_update_cls_tower_device.restype = c_int
    lib.sylph_update_cls_tower_device.argtypes = [vp, POINTER(TowerTensors), vp]
    lib.sylph_debug_read_buffer.restype = c_int
    lib.sylph_debug_read_buffer.argtypes = [vp, c_char_p, vp, ctypes.c_size_t, vp]
    lib.sylph_launch_count.restype = c_int64
    lib.sylph_launch_count.argtypes = [vp]
    lib.sylph_set_profiling.restype = c_int
    lib.sylph_set_profiling.argtypes = [vp, c_int]
    lib.sylph_get_timings.restype = c_int
    lib.sylph_get_timings.argtypes = [vp, vp, fp, POINTER(c_double), POINTER(c_double), c_int]
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "sylph_version", "sylph_create", "sylph_destroy", "sylph_last_error", "sylph_set_precision", "sylph_get_precision", "sylph_load_tensor",
    "sylph_finalize_weights", "sylph_extract_features", "sylph_extract_features_u8", "sylph_extract_features_normalized", "sylph_extract_features_multi", "sylph_import_features", "sylph_set_image_sizes", "sylph_feature_shape",
    "sylph_export_features", "sylph_generate_codes", "sylph_export_roi_features", "sylph_normalize_codes", "sylph_exchange_create", "sylph_exchange_connect",
    "sylph_normalize_codes_exchange", "sylph_exchange_poll", "sylph_exchange_status", "sylph_exchange_destroy", "sylph_accumulate_codes", "sylph_reduce_codes",
    "sylph_detect", "sylph_detect_after", "sylph_detect_poll", "sylph_export_head_output", "sylph_fcos_loss_sums", "sylph_fcos_loss_finalize",
    "sylph_fcos_cls_loss_backward", "sylph_codegen_backward", "sylph_update_code_generator", "sylph_update_code_generator_device", "sylph_set_training", "sylph_set_loss_box_branch", "sylph_cls_tower_backward",
    "sylph_update_cls_tower_device", "sylph_debug_read_buffer", "sylph_launch_count", "sylph_set_profiling", "sylph_get_timings",
]
