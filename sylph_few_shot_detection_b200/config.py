"""Config node + defaults for the Meta-FCOS inference path.

The reference assembles its defaults as d2go defaults -> vendored AdelaiDet FCOS keys -> Sylph adders
(`sylph/runner/meta_fcos_runner.py:104-114`, `sylph/runner/adet_configs.py:12-61`,
`sylph/runner/default_configs.py:9-160`) and reads YAML files through a yacs-style `CfgNode` with `_BASE_`
inheritance and `sylph://` rerouting (`sylph/config/config.py:20-65`).  Neither yacs, detectron2 nor d2go is
installed here, so this module restates the container and every default the hot path reads (SURVEY.md section 5,
"Config / flags").  Reference YAMLs (`configs/**/Meta-FCOS-*.yaml`) load unchanged; keys that belong to
subsystems outside the hot path (SOLVER, D2GO_DATA, DATALOADER ...) are accepted and stored verbatim.
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Dict, Iterable, List, Optional

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    """Attribute-access dict tree with yacs-compatible merge semantics (the subset the reference uses)."""

    def __init__(self, init: Optional[Dict[str, Any]] = None):
        super().__init__()
        if init:
            for k, v in init.items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name: str) -> Any:
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    # yacs API no-ops the reference calls
    def defrost(self) -> None:
        pass

    def freeze(self) -> None:
        pass

    def is_frozen(self) -> bool:
        return False

    def dump(self, **kwargs) -> str:
        def to_plain(n):
            if isinstance(n, dict):
                return {k: to_plain(v) for k, v in n.items()}
            if isinstance(n, tuple):
                return list(n)
            return n

        return yaml.safe_dump(to_plain(self), **kwargs)

    # ------------------------------------------------------------------ merging
    @staticmethod
    def _decode(v: Any) -> Any:
        """yacs decodes strings such as "(640, 672)" into Python literals."""
        if isinstance(v, str):
            try:
                return ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return v

    def merge_from_other_cfg(self, other: Dict[str, Any]) -> None:
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = self._decode(copy.deepcopy(v))

    @classmethod
    def load_yaml_with_base(cls, filename: str) -> Dict[str, Any]:
        filename = reroute_config_path(filename)
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}
        if BASE_KEY in cfg:
            base = reroute_config_path(cfg.pop(BASE_KEY))
            if not os.path.isabs(base) and not base.startswith("~"):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = cls.load_yaml_with_base(base)

            def merge_a_into_b(a, b):
                for k, v in a.items():
                    if isinstance(v, dict) and isinstance(b.get(k), dict):
                        merge_a_into_b(v, b[k])
                    else:
                        b[k] = v

            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename: str) -> None:
        self.merge_from_other_cfg(self.load_yaml_with_base(cfg_filename))

    def merge_from_list(self, opts: Iterable[Any]) -> None:
        opts = list(opts)
        assert len(opts) % 2 == 0, "override list must be KEY VALUE pairs"
        for key, value in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    node[p] = CfgNode()
                node = node[p]
            node[parts[-1]] = self._decode(value)


_CONFIG_ROOT = os.environ.get("SYLPH_CONFIG_ROOT", "")


def reroute_config_path(path: str) -> str:
    """`sylph://X` -> `<config root>/X` (reference: `sylph/config/config.py:33-44`).  The root is the directory
    holding the reference's `configs/` tree; set it with `set_config_root` or $SYLPH_CONFIG_ROOT."""
    if path.startswith("sylph://"):
        if not _CONFIG_ROOT:
            raise FileNotFoundError("sylph:// config path used but no config root set (set_config_root)")
        return os.path.join(_CONFIG_ROOT, path[len("sylph://"):])
    return path


def set_config_root(root: str) -> None:
    global _CONFIG_ROOT
    _CONFIG_ROOT = root


CN = CfgNode


def _detectron2_subset() -> CfgNode:
    """detectron2 defaults for the keys the hot path reads (upstream `detectron2/config/defaults.py`, restated)."""
    _C = CN()
    _C.VERSION = 2
    _C.MODEL = CN()
    _C.MODEL.DEVICE = "cuda"
    _C.MODEL.META_ARCHITECTURE = "GeneralizedRCNN"
    _C.MODEL.WEIGHTS = ""
    _C.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]
    _C.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]
    _C.MODEL.BACKBONE = CN()
    _C.MODEL.BACKBONE.NAME = "build_resnet_backbone"
    _C.MODEL.BACKBONE.FREEZE_AT = 2
    _C.MODEL.FPN = CN()
    _C.MODEL.FPN.IN_FEATURES = []
    _C.MODEL.FPN.OUT_CHANNELS = 256
    _C.MODEL.FPN.NORM = ""
    _C.MODEL.FPN.FUSE_TYPE = "sum"
    _C.MODEL.PROPOSAL_GENERATOR = CN()
    _C.MODEL.PROPOSAL_GENERATOR.NAME = "RPN"
    _C.MODEL.PROPOSAL_GENERATOR.MIN_SIZE = 0
    _C.MODEL.ROI_HEADS = CN()
    _C.MODEL.ROI_BOX_HEAD = CN()
    _C.MODEL.RESNETS = CN()
    _C.MODEL.RESNETS.DEPTH = 50
    _C.MODEL.RESNETS.OUT_FEATURES = ["res4"]
    _C.MODEL.RESNETS.NUM_GROUPS = 1
    _C.MODEL.RESNETS.NORM = "FrozenBN"
    _C.MODEL.RESNETS.WIDTH_PER_GROUP = 64
    _C.MODEL.RESNETS.STRIDE_IN_1X1 = True
    _C.MODEL.RESNETS.RES5_DILATION = 1
    _C.MODEL.RESNETS.RES2_OUT_CHANNELS = 256
    _C.MODEL.RESNETS.STEM_OUT_CHANNELS = 64
    _C.MODEL.RESNETS.DEFORM_ON_PER_STAGE = [False, False, False, False]
    _C.MODEL.RESNETS.DEFORM_MODULATED = False
    _C.MODEL.RESNETS.DEFORM_NUM_GROUPS = 1
    _C.INPUT = CN()
    _C.INPUT.MIN_SIZE_TRAIN = (800,)
    _C.INPUT.MIN_SIZE_TEST = 800
    _C.INPUT.MAX_SIZE_TEST = 1333
    _C.INPUT.FORMAT = "BGR"
    _C.INPUT.CROP = CN()
    _C.DATASETS = CN()
    _C.DATASETS.TRAIN = ()
    _C.DATASETS.TEST = ()
    _C.DATALOADER = CN()
    _C.SOLVER = CN()
    _C.TEST = CN()
    _C.TEST.DETECTIONS_PER_IMAGE = 100
    _C.SEED = -1
    return _C


def add_adet_fcos_config(_C: CfgNode) -> CfgNode:
    """FCOS keys; values follow the reference's vendored copy, `sylph/runner/adet_configs.py:25-61`."""
    _C.MODEL.MOBILENET = False
    _C.MODEL.BACKBONE.ANTI_ALIAS = False
    _C.MODEL.RESNETS.DEFORM_INTERVAL = 1
    _C.MODEL.FCOS = CN()
    F = _C.MODEL.FCOS
    F.NUM_CLASSES = 80
    F.IN_FEATURES = ["p3", "p4", "p5", "p6", "p7"]
    F.FPN_STRIDES = [8, 16, 32, 64, 128]
    F.PRIOR_PROB = 0.01
    F.INFERENCE_TH_TRAIN = 0.05
    F.INFERENCE_TH_TEST = 0.05
    F.NMS_TH = 0.6
    F.PRE_NMS_TOPK_TRAIN = 1000
    F.PRE_NMS_TOPK_TEST = 1000
    F.POST_NMS_TOPK_TRAIN = 100
    F.POST_NMS_TOPK_TEST = 100
    F.TOP_LEVELS = 2
    F.NORM = "GN"
    F.USE_SCALE = True
    F.THRESH_WITH_CTR = False
    F.LOSS_ALPHA = 0.25
    F.LOSS_GAMMA = 2.0
    F.SIZES_OF_INTEREST = [64, 128, 256, 512]
    F.USE_RELU = True
    F.USE_DEFORMABLE = False
    F.NUM_CLS_CONVS = 4
    F.NUM_BOX_CONVS = 4
    F.NUM_SHARE_CONVS = 0
    F.CENTER_SAMPLE = True
    F.POS_RADIUS = 1.5
    F.LOC_LOSS_TYPE = "giou"
    F.YIELD_PROPOSAL = False
    return _C


def add_sylph_config(_C: CfgNode) -> CfgNode:
    """Sylph adders, `sylph/runner/default_configs.py:9-160` (only keys are restated, in the reference's order)."""
    # add_base_config :9-41
    _C.DATASETS.ID_TRAIN = [0]
    _C.DATASETS.ID_TEST = [0]
    _C.DATASETS.BASE_CLASSES_SPLIT = ""
    _C.DATASETS.NOVEL_CLASSES_SPLIT = ""
    _C.DATASETS.NUMS_CLASSES = [0]
    _C.MODEL.WEIGHTS_FILTER_BY_MODULE = []
    _C.TEST.EVAL_PERIOD = 0
    _C.MODEL.BACKBONE.FREEZE = False
    _C.MODEL.BACKBONE.FREEZE_EXCLUDE = []
    P = _C.MODEL.PROPOSAL_GENERATOR
    P.OWD = False
    P.FREEZE_CLS_TOWER = False
    P.FREEZE_CLS_LOGITS = False
    P.FREEZE_BBOX_BRANCH = False
    P.FREEZE_BBOX_TOWER = False
    P.FREEZE = False
    _C.MODEL.ROI_HEADS.FREEZE = False
    _C.SEED = -1
    # add_fcos_config :44-50
    _C.MODEL.FCOS.BOX_QUALITY = ["ctrness"]
    _C.MODEL.FCOS.IOU_MASK = False
    _C.MODEL.FCOS.CLS_LOGITS_KERNEL_SIZE = 1
    _C.MODEL.FCOS.L2_NORM_CLS_WEIGHT = False
    # add_tfa_config :53-62
    _C.MODEL.TFA = CN()
    _C.MODEL.TFA.FINETINE = False
    _C.MODEL.TFA.TRAIN_SHOT = 10
    _C.MODEL.TFA.USE_PRETRAINED_BASE_CLS_LOGITS = True
    _C.MODEL.TFA.EVAL_WITH_PRETRAINED_BASE_CLS_LOGITS = False
    # add_default_meta_learn_config :65-96
    _C.MODEL.META_LEARN = CN()
    M = _C.MODEL.META_LEARN
    M.EPISODIC_LEARNING = False
    M.SHOT = 5
    M.EVAL_SHOT = 10
    M.BASE_EVAL_SHOT = 10
    M.CLASS = 5
    M.USE_ALL_GTS_IN_BASE_CLASSES = True
    M.EVAL_WITH_PRETRAINED_CODE = False
    M.QUERY_SHOT = 1
    M.CODE_GENERATOR = CN()
    G = M.CODE_GENERATOR
    G.FREEZE = False
    G.DISTILLATION_LOSS_WEIGHT = 0.0
    G.NAME = "CodeGenerator"
    G.ROI_BOX = CN()
    G.ROI_BOX.POOLER_RESOLUTION = 7
    G.ROI_BOX.POOLER_TYPE = "ROIAlignV2"
    G.ROI_BOX.FPN_MULTILEVEL_FEATURE = False
    _C.TEST.REPEAT_TEST = 1
    # add_code_genertor_config :99-140
    G.USE_MASK = True
    G.ALL_MASK = False
    G.MASK_NORM = "GN"
    G.CONV_L2_NORM = False
    G.USE_BIAS = True
    G.BIAS_L2_NORM = False
    G.TOWER_LAYERS = [["GN", ""]]
    G.CLS_LAYER = ["GN", "", 1]
    G.USE_WEIGHT_SCALE = True
    G.BIAS_LAYER = []
    G.WEIGHT_LAYER = []
    G.SCALE_LAYER = []
    G.BOX_ON = False
    G.BOX_TOWER_LAYERS = []
    G.BOX_CLS_LAYER = ["", "", 2]
    G.BOX_BIAS_LAYER = []
    G.CONTRASTIVE_LOSS = ""
    G.INIT_NORM_LAYER = False
    G.CLS_REWEIGHT = False
    G.META_WEIGHT = False
    G.META_BIAS = False
    G.USE_PER_CLS_SCALE = False
    G.COMPRESS_CODE_W_MAX = False
    G.POST_NORM = "GN"
    G.IN_CHANNEL = 256
    G.OUT_CHANNEL = 256
    G.USE_DEFORMABLE = False
    # add_roi_encoder_config :143-160
    G.TOKENIZER = CN()
    G.TOKENIZER.NUM_CONV = 0
    G.TOKENIZER.CONV_DIM = 256
    G.TOKENIZER.NORM = ""
    G.TOKENIZER.NUM_FC = 1
    G.TOKENIZER.FC_DIM = 256
    G.TRANSFORMER_ENCODER = CN()
    G.TRANSFORMER_ENCODER.LAYERS = 1
    G.TRANSFORMER_ENCODER.HEADS = 8
    G.TRANSFORMER_ENCODER.DROPOUT = 0.1
    G.HEAD = CN()
    G.HEAD.NUM_FC = 1
    G.HEAD.FC_DIM = 512
    G.HEAD.OUTPUT_DIM = 256
    return _C


def get_default_cfg() -> CfgNode:
    """Equivalent of `MetaFCOSRunner.get_default_cfg()` (`sylph/runner/meta_fcos_runner.py:104-114`)."""
    return add_sylph_config(add_adet_fcos_config(_detectron2_subset()))


def load_cfg(config_file: str, opts: Optional[List[Any]] = None) -> CfgNode:
    cfg = get_default_cfg()
    cfg.merge_from_file(config_file)
    if opts:
        cfg.merge_from_list(opts)
    return cfg
