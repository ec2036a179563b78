"""Configs of the golden cases, restated as overrides on top of the defaults (the GPU box has no /root/reference,
so the reference YAML files cannot be read there).  tests/test_config.py checks, in the build container, that these
overrides reproduce the hot-path keys of the reference YAMLs."""
import os

import torch

from sylph_few_shot_detection_b200.config import get_default_cfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_COMMON = [
    "MODEL.META_ARCHITECTURE", "MetaOneStageDetector",
    "MODEL.BACKBONE.NAME", "build_fcos_resnet_fpn_backbone",
    "MODEL.BACKBONE.FREEZE", True,
    "MODEL.RESNETS.OUT_FEATURES", ["res3", "res4", "res5"],
    "MODEL.RESNETS.DEPTH", 50,
    "MODEL.FPN.IN_FEATURES", ["res3", "res4", "res5"],
    "MODEL.PROPOSAL_GENERATOR.NAME", "MetaFCOS",
    "MODEL.PROPOSAL_GENERATOR.FREEZE_BBOX_BRANCH", True,
    "MODEL.META_LEARN.EPISODIC_LEARNING", True,
    "MODEL.META_LEARN.USE_ALL_GTS_IN_BASE_CLASSES", False,
    "MODEL.META_LEARN.CLASS", 3,
    "MODEL.META_LEARN.CODE_GENERATOR.CONV_L2_NORM", True,
    "MODEL.META_LEARN.CODE_GENERATOR.TOWER_LAYERS", [["GN", "ReLU"], ["GN", "ReLU"]],
    "MODEL.META_LEARN.CODE_GENERATOR.CLS_LAYER", ["", "", 1],
    "MODEL.META_LEARN.CODE_GENERATOR.BIAS_LAYER", ["", "", 1],
    "MODEL.FCOS.BOX_QUALITY", ["ctrness"],
    "MODEL.DEVICE", "cpu",
]
OVERRIDES = {
    # configs/COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml (+ Base-FCOS.yaml)
    "COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml": _COMMON + [
        "MODEL.FCOS.NUM_CLASSES", 60,
    ],
    # configs/LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml (+ Base-Meta-FCOS.yaml)
    "LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml": _COMMON + [
        "MODEL.FCOS.NUM_CLASSES", 866,
        "MODEL.FCOS.POST_NMS_TOPK_TEST", 300,
        "MODEL.FCOS.POST_NMS_TOPK_TRAIN", 300,
        "MODEL.META_LEARN.CODE_GENERATOR.BIAS_L2_NORM", True,
        "MODEL.META_LEARN.CODE_GENERATOR.USE_PER_CLS_SCALE", True,
        "MODEL.TFA.USE_PRETRAINED_BASE_CLS_LOGITS", False,
    ],
}


def cfg_for(config_name: str):
    cfg = get_default_cfg()
    cfg.merge_from_list(OVERRIDES[config_name])
    return cfg


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| relative to max |b| (the tensor scale)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0 if a.numel() == 0 else float("inf")
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
