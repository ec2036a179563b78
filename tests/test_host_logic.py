"""Host-side mirror of the plugin surface: registry names, run_type protocol, error behaviour, code packing --
everything that can be checked without a device."""
import numpy as np
import pytest
import torch

from sylph_few_shot_detection_b200 import modeling as M
from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
from sylph_few_shot_detection_b200.runner import MetaFCOSRunner, format_class_codes_shared, shard_range
from sylph_few_shot_detection_b200.structures import Boxes, Instances


def test_registries_expose_the_reference_names():
    assert M.CODE_GENERATOR_REGISTRY.get("CodeGenerator") is M.CodeGenerator
    assert M.CODE_GENERATOR_REGISTRY.get("CodeGeneratorHead") is M.CodeGenerator
    assert M.META_ARCH_REGISTRY.get("MetaOneStageDetector") is M.MetaOneStageDetector
    assert M.PROPOSAL_GENERATOR_REGISTRY.get("MetaFCOS") is M.MetaFCOS
    assert M.BACKBONE_REGISTRY.get("build_fcos_resnet_fpn_backbone") is M.build_fcos_resnet_fpn_backbone
    with pytest.raises(KeyError):
        M.CODE_GENERATOR_REGISTRY.get("NoSuchGenerator")


def test_model_tree_and_run_type_protocol():
    cfg = coco_meta_fcos_cfg()
    model = MetaFCOSRunner().build_model(cfg)
    assert not model.training and model.episodic_learning
    shapes = model.backbone.output_shape()
    assert [shapes[f].stride for f in cfg.MODEL.FCOS.IN_FEATURES] == [8, 16, 32, 64, 128]
    assert model.backbone.size_divisibility == 32
    for name in ("cls_tower", "bbox_tower", "share_tower", "cls_logits", "bbox_pred", "ctrness", "iou_overlap"):
        assert hasattr(model.proposal_generator.fcos_head, name)
    assert list(model.code_generator.parameters()) == []
    with pytest.raises(NotImplementedError):
        model([], run_type="no_such_type")
    with pytest.raises(AssertionError):
        model([{"support_set": []}, {"support_set": []}], run_type="meta_learn_test_support")
    with pytest.raises(RuntimeError):  # no weights / no device: loud failure, no CPU fallback
        model.engine


def test_load_state_dict_validates_keys_and_shapes():
    from sylph_few_shot_detection_b200 import weights as W
    cfg = coco_meta_fcos_cfg()
    model = M.build_model(cfg)
    spec = W.state_spec(cfg)
    state = {k: torch.zeros(s) for k, s in spec.items()}
    bad = dict(state)
    del bad["backbone.fpn_lateral3.weight"]
    with pytest.raises(RuntimeError, match="missing"):
        model.load_state_dict(bad)
    bad = dict(state)
    bad["backbone.fpn_lateral3.weight"] = torch.zeros(1, 2)
    with pytest.raises(RuntimeError, match="size mismatch"):
        model.load_state_dict(bad)


def test_select_a_mask_consumes_the_global_numpy_rng_and_rejects_empty():
    inst = Instances((10, 10))
    inst.gt_boxes = Boxes(torch.tensor([[0.0, 0.0, 1.0, 1.0], [1.0, 1.0, 2.0, 2.0], [2.0, 2.0, 3.0, 3.0]]))
    np.random.seed(7)
    expect = np.random.choice(range(3), 1)
    np.random.seed(7)
    got = M.select_a_mask([inst])
    assert torch.equal(got[0], inst.gt_boxes.tensor[expect])
    empty = Instances((10, 10))
    empty.gt_boxes = Boxes(torch.zeros(0, 4))
    with pytest.raises(ValueError):
        M.select_a_mask([empty])


def test_all_mask_takes_every_box_and_keeps_the_reference_shape_check():
    """CODE_GENERATOR.ALL_MASK (utils.py:27-47, code_generator.py:928-937): no RNG draw; the pooler must still return one
    ROI per support image, so an image with two boxes trips the reference's own assertion."""
    def inst(n):
        i = Instances((64, 64))
        i.gt_boxes = Boxes(torch.arange(4.0 * n).reshape(n, 4) + 1.0)
        return i
    np.random.seed(3)
    state = np.random.get_state()[1].copy()
    got = M.support_boxes([inst(1), inst(1)], True, 2)
    assert got.shape == (2, 4) and np.array_equal(np.random.get_state()[1], state)       # global RNG untouched
    with pytest.raises(AssertionError, match="pooled_features.shape"):
        M.support_boxes([inst(2), inst(1)], True, 2)
    assert M.support_boxes([inst(3), inst(2)], False, 2).shape == (2, 4)                   # one drawn box per image


def test_pack_and_format_class_codes():
    codes = [{"support_set_target": torch.tensor(1), "class_code": {"cls_conv": torch.full((1, 256, 1, 1), 2.0), "cls_bias": torch.tensor([0.2])}},
             {"support_set_target": 0, "class_code": {"cls_conv": torch.full((1, 256, 1, 1), 0.5), "cls_bias": torch.tensor([0.0])}}]
    packed = format_class_codes_shared(codes)
    assert packed["cls_conv"].shape == (2, 256, 1, 1) and packed["cls_bias"].shape == (2,)
    assert float(packed["cls_conv"][0, 0, 0, 0]) == 0.5 and abs(float(packed["cls_bias"][1]) - 0.2) < 1e-7
    rows = M.pack_code_rows(packed)
    assert rows.shape == (2, 257) and float(rows[1, 256]) == pytest.approx(0.2)
    with pytest.raises(AssertionError):
        M.pack_code_rows({"cls_conv": torch.zeros(2, 128, 1, 1), "cls_bias": torch.zeros(2)})


def test_shard_range_is_contiguous_balanced_and_complete():
    for n, w in [(20, 8), (5, 8), (1203, 8), (8, 2), (0, 4)]:
        parts = [list(shard_range(n, w, r)) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [len(shard_range(20, 8, r)) for r in range(8)] == [3, 3, 3, 3, 2, 2, 2, 2]


def test_balanced_query_assignment_fills_the_lightest_ranks_first():
    from sylph_few_shot_detection_b200.runner import balanced_query_assignment, query_indices_of_rank
    # BASELINE configs[3]: 20 classes x 5 shots on 8 ranks (15,15,15,15,10,10,10,10 support images), 8 query images
    a = balanced_query_assignment([15, 15, 15, 15, 10, 10, 10, 10], 8)
    assert a == [[], [], [], [], [0, 4], [1, 5], [2, 6], [3, 7]]
    assert sorted(q for r in a for q in r) == list(range(8))
    # equal loads: round-robin from rank 0, i.e. one image per rank like the contiguous shards
    assert balanced_query_assignment([5, 5, 5, 5], 4) == [[0], [1], [2], [3]]
    assert balanced_query_assignment([3], 2) == [[0, 1]]
    support = [{"support_set": [None] * 5} for _ in range(20)]
    got = [query_indices_of_rank(support, 8, 8, r, balance_queries=True) for r in range(8)]
    assert got == a
    assert [query_indices_of_rank(support, 8, 8, r) for r in range(8)] == [[r] for r in range(8)]
