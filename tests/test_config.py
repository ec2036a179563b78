"""Config container semantics and the presets against the reference YAMLs / model (container-only parts are marked
`reference`)."""
import os

import pytest
import torch

from sylph_few_shot_detection_b200 import weights as W
from sylph_few_shot_detection_b200.config import CfgNode, get_default_cfg, load_cfg, set_config_root
from sylph_few_shot_detection_b200.presets import OVERRIDES, preset_cfg

REF_CFG = "/root/reference/configs"
HOT_KEYS = ["PIXEL_MEAN", "PIXEL_STD", "FCOS", "META_LEARN", "RESNETS", "FPN", "BACKBONE", "PROPOSAL_GENERATOR", "TFA",
            "META_ARCHITECTURE"]


def test_defaults_hold_the_reference_values():
    c = get_default_cfg()
    assert c.MODEL.FCOS.INFERENCE_TH_TEST == 0.05 and c.MODEL.FCOS.PRE_NMS_TOPK_TEST == 1000
    assert c.MODEL.FCOS.POST_NMS_TOPK_TEST == 100 and c.MODEL.FCOS.NMS_TH == 0.6
    assert c.MODEL.FCOS.FPN_STRIDES == [8, 16, 32, 64, 128] and c.MODEL.FCOS.CLS_LOGITS_KERNEL_SIZE == 1
    assert c.MODEL.META_LEARN.CODE_GENERATOR.POST_NORM == "GN" and c.MODEL.META_LEARN.CODE_GENERATOR.USE_WEIGHT_SCALE
    assert c.MODEL.PIXEL_MEAN == [103.530, 116.280, 123.675]


def test_cfgnode_merge_list_base_and_literals(tmp_path):
    base = tmp_path / "base.yaml"
    base.write_text("MODEL:\n  FCOS:\n    NUM_CLASSES: 7\nINPUT:\n  MIN_SIZE_TRAIN: (640, 672)\n")
    child = tmp_path / "child.yaml"
    child.write_text('_BASE_: "base.yaml"\nMODEL:\n  FCOS:\n    NMS_TH: 0.5\nNEW_SECTION:\n  X: 1\n')
    c = get_default_cfg()
    c.merge_from_file(str(child))
    assert c.MODEL.FCOS.NUM_CLASSES == 7 and c.MODEL.FCOS.NMS_TH == 0.5 and c.NEW_SECTION.X == 1
    assert c.INPUT.MIN_SIZE_TRAIN == (640, 672)
    c.merge_from_list(["MODEL.FCOS.NMS_TH", 0.7, "MODEL.DEVICE", "cpu"])
    assert c.MODEL.FCOS.NMS_TH == 0.7 and c.MODEL.DEVICE == "cpu"
    d = c.clone()
    d.MODEL.FCOS.NMS_TH = 0.1
    assert c.MODEL.FCOS.NMS_TH == 0.7
    set_config_root(str(tmp_path))
    assert load_cfg("sylph://child.yaml").MODEL.FCOS.NUM_CLASSES == 7


def test_state_spec_and_synthetic_fill_are_deterministic():
    cfg = preset_cfg("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml")
    spec = W.state_spec(cfg)
    assert len(spec) == 354
    assert spec["code_generator.code_generator_head.support_set_cls_bias.0.weight"] == (1, 256, 3, 3)
    assert spec["proposal_generator.fcos_head.cls_logits.weight"] == (60, 256, 1, 1)
    a = W.synthetic_tensor(cfg, "backbone.fpn_output3.weight", spec["backbone.fpn_output3.weight"], 3)
    b = W.synthetic_tensor(cfg, "backbone.fpn_output3.weight", spec["backbone.fpn_output3.weight"], 3)
    c = W.synthetic_tensor(cfg, "backbone.fpn_output3.weight", spec["backbone.fpn_output3.weight"], 4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    cfg101 = preset_cfg("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", ["MODEL.RESNETS.DEPTH", 101])
    assert "backbone.bottom_up.res4.22.conv3.weight" in W.state_spec(cfg101)


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(OVERRIDES))
def test_presets_reproduce_the_reference_yaml_hot_path_keys(name):
    ref = load_cfg(os.path.join(REF_CFG, name))
    mine = preset_cfg(name)
    for k in HOT_KEYS:
        assert ref.MODEL[k] == mine.MODEL[k], k


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(OVERRIDES))
def test_state_spec_equals_reference_model_state_dict(name):
    import contextlib
    import io
    from oracle import reference_loader
    cfg = preset_cfg(name, ["MODEL.DEVICE", "cpu"])
    with contextlib.redirect_stdout(io.StringIO()):
        model = reference_loader.build_reference_model(cfg)
    ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ref == dict(W.state_spec(cfg))


def test_unsupported_training_variants_fail_loudly():
    """SURVEY K9 (soft-nearest-neighbour loss), per-class box regression and distillation are outside the path: the loss
    configuration refuses them instead of computing something else."""
    import pytest
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runtime import loss_config_from_cfg
    loss_config_from_cfg(coco_meta_fcos_cfg())
    for opts in (["MODEL.META_LEARN.CODE_GENERATOR.CONTRASTIVE_LOSS", "snnl"],
                 ["MODEL.META_LEARN.CODE_GENERATOR.BOX_ON", True],
                 ["MODEL.META_LEARN.CODE_GENERATOR.DISTILLATION_LOSS_WEIGHT", 0.5],
                 ["MODEL.FCOS.LOC_LOSS_TYPE", "diou"]):
        with pytest.raises(NotImplementedError):
            loss_config_from_cfg(coco_meta_fcos_cfg(opts))
