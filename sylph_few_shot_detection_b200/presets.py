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
    # base detector (pre-training model, no code generator): `run_type=None` inference with the model's own cls_logits
    "COCO-Detection/Meta-FCOS/Meta-FCOS-pretrain.yaml": [
        "MODEL.META_ARCHITECTURE", "MetaOneStageDetector",
        "MODEL.BACKBONE.NAME", "build_fcos_resnet_fpn_backbone",
        "MODEL.BACKBONE.FREEZE", False,
        "MODEL.RESNETS.OUT_FEATURES", ["res3", "res4", "res5"],
        "MODEL.RESNETS.DEPTH", 50,
        "MODEL.FPN.IN_FEATURES", ["res3", "res4", "res5"],
        "MODEL.PROPOSAL_GENERATOR.NAME", "MetaFCOS",
        "MODEL.META_LEARN.EPISODIC_LEARNING", False,
        "MODEL.FCOS.NUM_CLASSES", 60,
        "MODEL.DDP_FIND_UNUSED_PARAMETERS", True,
    ],
    # ROIEncoder generator (transformer hyper-network); inherits Base-Meta-FCOS.yaml, not the CodeGenerator overrides
    "LVISv1-Detection/Meta-FCOS/Meta-FCOS-ROI-Encoder-finetune.yaml": [
        "MODEL.META_ARCHITECTURE", "MetaOneStageDetector",
        "MODEL.BACKBONE.NAME", "build_fcos_resnet_fpn_backbone",
        "MODEL.BACKBONE.FREEZE", True,
        "MODEL.RESNETS.OUT_FEATURES", ["res3", "res4", "res5"],
        "MODEL.RESNETS.DEPTH", 50,
        "MODEL.FPN.IN_FEATURES", ["res3", "res4", "res5"],
        "MODEL.PROPOSAL_GENERATOR.NAME", "MetaFCOS",
        "MODEL.PROPOSAL_GENERATOR.OWD", False,
        "MODEL.PROPOSAL_GENERATOR.FREEZE_CLS_TOWER", True,
        "MODEL.PROPOSAL_GENERATOR.FREEZE_BBOX_BRANCH", True,
        "MODEL.DDP_FIND_UNUSED_PARAMETERS", True,
        "MODEL.FCOS.NUM_CLASSES", 1103,
        "MODEL.FCOS.POST_NMS_TOPK_TEST", 300,
        "MODEL.FCOS.POST_NMS_TOPK_TRAIN", 300,
        "MODEL.FCOS.NUM_CLS_CONVS", 4,
        "MODEL.FCOS.CLS_LOGITS_KERNEL_SIZE", 1,
        "MODEL.FCOS.NORM", "GN",
        "MODEL.FCOS.BOX_QUALITY", ["ctrness"],
        "MODEL.FCOS.CENTER_SAMPLE", True,
        "MODEL.FCOS.IOU_MASK", False,
        "MODEL.META_LEARN.EPISODIC_LEARNING", True,
        "MODEL.META_LEARN.CODE_GENERATOR.NAME", "ROIEncoder",
        "MODEL.META_LEARN.CODE_GENERATOR.ROI_BOX.POOLER_RESOLUTION", 7,
        "MODEL.META_LEARN.CODE_GENERATOR.ROI_BOX.POOLER_TYPE", "ROIAlignV2",
        "MODEL.META_LEARN.CODE_GENERATOR.TOKENIZER.NUM_CONV", 2,
        "MODEL.META_LEARN.CODE_GENERATOR.TOKENIZER.CONV_DIM", 256,
        "MODEL.META_LEARN.CODE_GENERATOR.TOKENIZER.NORM", "GN",
        "MODEL.META_LEARN.CODE_GENERATOR.TOKENIZER.NUM_FC", 2,
        "MODEL.META_LEARN.CODE_GENERATOR.TOKENIZER.FC_DIM", 256,
        "MODEL.META_LEARN.CODE_GENERATOR.TRANSFORMER_ENCODER.LAYERS", 2,
        "MODEL.META_LEARN.CODE_GENERATOR.TRANSFORMER_ENCODER.HEADS", 8,
        "MODEL.META_LEARN.CODE_GENERATOR.TRANSFORMER_ENCODER.DROPOUT", 0.1,
        "MODEL.META_LEARN.CODE_GENERATOR.HEAD.NUM_FC", 2,
        "MODEL.META_LEARN.CODE_GENERATOR.HEAD.FC_DIM", 512,
        "MODEL.META_LEARN.CODE_GENERATOR.HEAD.OUTPUT_DIM", 256,
        "MODEL.META_LEARN.CLASS", 3,
        "MODEL.META_LEARN.SHOT", 5,
        "MODEL.META_LEARN.EVAL_SHOT", 10,
        "MODEL.META_LEARN.QUERY_SHOT", 1,
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


def lvis_roi_encoder_cfg(opts: Optional[List[Any]] = None) -> CfgNode:
    return preset_cfg("LVISv1-Detection/Meta-FCOS/Meta-FCOS-ROI-Encoder-finetune.yaml", opts)


def lvis_meta_fcos_cfg(opts: Optional[List[Any]] = None) -> CfgNode:
    return preset_cfg("LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", opts)
