"""The shipped Meta-FCOS configs restated as overrides on top of the defaults, for environments where the reference's
YAML tree is not on disk (benchmarks, the GPU box).  tests/test_config.py checks them against the reference YAMLs
(`configs/COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml`, `configs/LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml`)
whenever /root/reference is present."""
from __future__ import annotations

from typing import Any, List, Optional

from .config import CfgNode, get_default_cfg

_COMMON: List[Any] = [
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
]

OVERRIDES = {
    "COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml": _COMMON + [
        "MODEL.FCOS.NUM_CLASSES", 60,
    ],
    "LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml": _COMMON + [
        "MODEL.FCOS.NUM_CLASSES", 866,
        "MODEL.FCOS.POST_NMS_TOPK_TEST", 300,
        "MODEL.FCOS.POST_NMS_TOPK_TRAIN", 300,
        "MODEL.META_LEARN.CODE_GENERATOR.BIAS_L2_NORM", True,
        "MODEL.META_LEARN.CODE_GENERATOR.USE_PER_CLS_SCALE", True,
        "MODEL.TFA.USE_PRETRAINED_BASE_CLS_LOGITS", False,
    ],
}


def preset_cfg(config_name: str, opts: Optional[List[Any]] = None) -> CfgNode:
    cfg = get_default_cfg()
    cfg.merge_from_list(OVERRIDES[config_name])
    if opts:
        cfg.merge_from_list(opts)
    return cfg


def coco_meta_fcos_cfg(opts: Optional[List[Any]] = None) -> CfgNode:
    return preset_cfg("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", opts)


def lvis_meta_fcos_cfg(opts: Optional[List[Any]] = None) -> CfgNode:
    return preset_cfg("LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", opts)
