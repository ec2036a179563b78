"""Persistent class-code store and single-image predictor ("registered classes") over the B200 inference path.

Mirrors the reference's
  * on-disk class-code format .... sylph/evaluation/meta_learn_evaluation.py:305-326: one `<class_name>.pth` per class
                                   = torch.save({"support_set_target", "class_name", ..., "class_code": {"cls_conv"
                                   (1,256,1,1), "cls_bias" (1,1,1,1)}}) with CPU tensors (RAW codes: they are written
                                   before inference_normalization runs);
  * `SylphPredictor` ............. sylph/predictor.py:37-298: codes are read back per class name
                                   (`_get_datasets_class_codes` :167-187, missing file -> ValueError), packed with
                                   format_class_codes_shared and fed to run_type="meta_learn_test_instance" for one
                                   BGR HxWx3 image after ResizeShortestEdge (`_call_few_shot` :248-274).
Dataset catalogs, visualisation and the d2go runner factory of the reference predictor are out of scope; class names are
passed in directly.  Files written by the reference load here and vice versa.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .modeling import MetaOneStageDetector, build_model
from .runner import _target_id, format_class_codes_shared, inference_normalization, inference_on_support_set


def save_class_codes(results: Sequence[Dict[str, Any]], output_dir: str) -> List[str]:
    """Write one `<class_name>.pth` per entry (meta_learn_evaluation.py:316-325): every key of the item except
    "support_set", `class_code` tensors moved to the CPU; an existing file is replaced."""
    os.makedirs(output_dir, exist_ok=True)
    paths = []
    for item in results:
        rec = {k: v for k, v in item.items() if k not in ("support_set", "class_code")}
        rec["class_code"] = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in item["class_code"].items()}
        path = os.path.join(output_dir, f"{rec['class_name']}.pth")
        if os.path.isfile(path):
            os.remove(path)
        with open(path, "wb") as f:
            torch.save(rec, f)
        paths.append(path)
    return paths


def load_class_code_list(code_dir: str, class_names: Sequence[str]) -> List[Dict[str, Any]]:
    """Read `<class_name>.pth` for every name, in order (predictor.py:173-183); a missing file raises ValueError."""
    codes = []
    for name in class_names:
        path = os.path.join(code_dir, f"{name}.pth")
        if not os.path.exists(path):
            raise ValueError(f"{path} is missing")
        with open(path, "rb") as f:
            codes.append(torch.load(f, map_location="cpu", weights_only=False))
    return codes


def resize_shortest_edge_shape(h: int, w: int, short: int, max_size: int):
    """detectron2 ResizeShortestEdge.get_output_shape (un-vendored upstream)."""
    scale = short * 1.0 / min(h, w)
    newh, neww = (short, scale * w) if h < w else (scale * h, short)
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def resize_image(img: np.ndarray, newh: int, neww: int) -> np.ndarray:
    """uint8 HxWxC bilinear resize through PIL, as detectron2's ResizeTransform.apply_image does for uint8 input."""
    from PIL import Image
    if img.shape[0] == newh and img.shape[1] == neww:
        return img
    return np.asarray(Image.fromarray(img).resize((neww, newh), Image.BILINEAR))


class SylphPredictor:
    """Single-image few-shot predictor over registered class codes (reference: sylph/predictor.py:37-298).

        pred = SylphPredictor(cfg, state_dict, class_code_path, {"all": ("lvis_val", class_names)})
        out = pred._call_few_shot(bgr_image, pred.class_codes["all"])["instances"]

    `normalize_codes=False` reproduces the reference exactly: the stored codes are RAW and the reference predictor feeds
    them to the head without `inference_normalization` (predictor.py:167-187 -> :248-274).  Pass True to apply the code
    normalisation the meta-test runner applies (meta_fcos_runner.py:538) before use.
    """

    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], class_code_path: Optional[str] = None,
                 test_datasets: Optional[Dict[str, Any]] = None, normalize_codes: bool = False):
        assert cfg.MODEL.META_LEARN.EPISODIC_LEARNING, "This is not few-shot model"
        self.cfg = cfg
        self.model: MetaOneStageDetector = build_model(cfg)
        self.model.load_state_dict(state_dict)
        self.class_code_path = class_code_path
        self.normalize_codes = normalize_codes
        self.input_format = cfg.INPUT.FORMAT
        assert self.input_format in ["RGB", "BGR"], self.input_format
        self.min_size, self.max_size = int(cfg.INPUT.MIN_SIZE_TEST), int(cfg.INPUT.MAX_SIZE_TEST)
        self.class_names: Dict[str, List[str]] = {}
        self.class_codes: Dict[str, Dict[str, torch.Tensor]] = {}
        self._registered: List[Dict[str, Any]] = []
        for split, (dataset_name, names) in (test_datasets or {}).items():
            self.class_names[split] = list(names)
            self.class_codes[split] = self._get_datasets_class_codes(names, dataset_name)

    # ------------------------------------------------------------------ code store
    def _pack(self, codes: List[Dict[str, Any]]) -> Dict[str, torch.Tensor]:
        codes = [dict(c, class_code=dict(c["class_code"])) for c in codes]
        if self.normalize_codes:
            codes = inference_normalization(self.model, codes)
        return format_class_codes_shared(codes, device=self.model.device)

    def _get_datasets_class_codes(self, class_names: Sequence[str], dataset_name: str, seed: int = 0):
        """predictor.py:167-187: `<class_code_path>/<dataset_name>/<seed>/<class_name>.pth`."""
        code_path = os.path.join(self.class_code_path, dataset_name, str(seed))
        codes = load_class_code_list(code_path, class_names)
        packed = self._pack(codes)
        assert "cls_conv" in packed, "conv is not in class_codes"
        return packed

    def generate_class_codes(self, support_items: Sequence[Dict[str, Any]], output_dir: Optional[str] = None):
        """`_generate_class_code_from_dataset` (predictor.py:131-161) for in-memory support sets: run the support pass,
        optionally persist the raw codes, return the packed codes."""
        results = inference_on_support_set(self.model, support_items)
        if output_dir is not None:
            save_class_codes(results, output_dir)
        return self._pack(results)

    def register_class(self, support_item: Dict[str, Any]) -> int:
        """Incremental class registration (the reference leaves `_generate_class_codes_from_a_support_set` unimplemented,
        predictor.py:163-165): one support set -> one more class in the "user" split, without touching the others."""
        res = inference_on_support_set(self.model, [support_item])[0]
        res["support_set_target"] = len(self._registered)
        self._registered.append(res)
        self.class_names["user"] = [r.get("class_name", "") for r in self._registered]
        self.class_codes["user"] = self._pack(self._registered)
        return _target_id(res["support_set_target"])

    # ------------------------------------------------------------------ inference
    def _call_few_shot(self, original_image: np.ndarray, class_codes: Dict[str, torch.Tensor]):
        """predictor.py:248-274: (H, W, 3) BGR uint8 -> {"instances": Instances} in the ORIGINAL image frame."""
        with torch.no_grad():
            if self.input_format == "RGB":
                original_image = original_image[:, :, ::-1]
            height, width = original_image.shape[:2]
            newh, neww = resize_shortest_edge_shape(height, width, self.min_size, self.max_size)
            image = resize_image(np.ascontiguousarray(original_image), newh, neww)
            image = torch.as_tensor(np.ascontiguousarray(image.transpose(2, 0, 1)))  # uint8 CHW: normalised in the stem
            inputs = {"image": image, "height": height, "width": width}
            return self.model([inputs], class_code=class_codes, run_type="meta_learn_test_instance")[0]

    def inference_on_registered_class(self, original_image: np.ndarray):
        return self._call_few_shot(original_image, self.class_codes["user"])

    def inference_on_split(self, original_image: np.ndarray, split: str = "all"):
        return self._call_few_shot(original_image, self.class_codes[split])

    def __call__(self, original_image: np.ndarray):
        raise NotImplementedError("base-detector inference (run_type=None) is not implemented on the B200 path; "
                                  "use _call_few_shot / inference_on_split with class codes")
