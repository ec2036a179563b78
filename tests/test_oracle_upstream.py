"""Pin the restated upstream operators (oracle/upstream.py) against the installed torchvision
ops and against hand-computed known answers."""
import math

import pytest
import torch

from oracle import upstream as up


def test_assign_boxes_to_levels_boundaries():
    # sqrt(area) s: s < 224 -> p3, [224, 448) -> p4, [448, 896) -> p5, [896, 1792) -> p6, >= 1792 -> p7
    sides = [8.0, 223.0, 224.0, 447.9, 448.0, 895.0, 896.0, 1791.0, 1792.0, 5000.0]
    boxes = [up.Boxes(torch.tensor([[0.0, 0.0, s, s]])) for s in sides]
    lv = up.assign_boxes_to_levels(boxes, 3, 7, 224, 4)
    assert lv.dtype == torch.int64
    assert lv.tolist() == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    # a full 1333x800 box has sqrt(area) ~ 1033 -> p6: p7 is never pooled from at this image size (Appendix A.4)
    assert int(up.assign_boxes_to_levels([up.Boxes(torch.tensor([[0.0, 0.0, 1333.0, 800.0]]))], 3, 7, 224, 4)) == 3


@pytest.mark.parametrize("box", [[3.0, 2.0, 20.0, 17.0], [0.0, 0.0, 31.0, 23.0], [-6.0, -4.0, 9.0, 40.0], [5.2, 5.1, 6.0, 5.9]])
def test_roi_align_restatement_matches_torchvision(box):
    from torchvision.ops import roi_align
    torch.manual_seed(0)
    x = torch.randn(2, 3, 24, 32)
    rois = torch.tensor([[1.0] + box])
    got = up.roi_align_restated(x, rois, 7, 0.5)
    ref = roi_align(x, rois, (7, 7), 0.5, 0, True)
    assert torch.allclose(got, ref, atol=2e-6, rtol=1e-5)


def test_roi_pooler_scatter_order_and_levels():
    torch.manual_seed(1)
    feats = [torch.randn(3, 4, 64 >> l, 96 >> l) for l in range(5)]
    boxes = [up.Boxes(torch.tensor([[10.0, 10.0, 60.0, 50.0]])), up.Boxes(torch.tensor([[0.0, 0.0, 500.0, 480.0]])),
             up.Boxes(torch.tensor([[100.0, 50.0, 400.0, 300.0]]))]
    pooler = up.ROIPooler(7, [1 / 8, 1 / 16, 1 / 32, 1 / 64, 1 / 128], 0, "ROIAlignV2")
    out = pooler(feats, boxes)
    assert out.shape == (3, 4, 7, 7)
    from torchvision.ops import roi_align
    lv = up.assign_boxes_to_levels(boxes, 3, 7, 224, 4).tolist()
    assert lv == [0, 2, 1]
    for i, l in enumerate(lv):
        roi = torch.cat([torch.tensor([[float(i)]]), boxes[i].tensor], dim=1)
        assert torch.equal(out[i], roi_align(feats[l], roi, (7, 7), 1.0 / (8 << l), 0, True)[0])


def test_nms_restatement_matches_torchvision():
    from torchvision.ops import nms
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(300, 2, generator=g) * 100
    wh = torch.rand(300, 2, generator=g) * 40 + 2
    boxes = torch.cat([xy, xy + wh], dim=1)
    scores = torch.rand(300, generator=g)
    assert torch.equal(up.nms_restated(boxes, scores, 0.6), nms(boxes, scores, 0.6))
    idxs = torch.randint(0, 4, (300,), generator=g)
    assert torch.equal(up.batched_nms(boxes, scores, idxs, 0.6, use_torchvision=False), up.batched_nms(boxes, scores, idxs, 0.6))
    assert up.batched_nms(boxes[:0], scores[:0], idxs[:0], 0.6).numel() == 0


def test_image_list_pads_after_normalisation_to_multiple_of_32():
    a, b = torch.ones(3, 50, 70), torch.ones(3, 64, 33) * 2
    il = up.ImageList.from_tensors([a, b], 32)
    assert il.tensor.shape == (2, 3, 64, 96) and il.image_sizes == [(50, 70), (64, 33)]
    assert float(il.tensor[0, :, 50:, :].abs().sum()) == 0 and float(il.tensor[1, :, :, 33:].abs().sum()) == 0
    assert float(il.tensor[1, 0, 63, 32]) == 2.0


def test_compute_locations_known_answer():
    loc = up.compute_locations(2, 3, 8, "cpu")
    assert loc.tolist() == [[4.0, 4.0], [12.0, 4.0], [20.0, 4.0], [4.0, 12.0], [12.0, 12.0], [20.0, 12.0]]


def test_frozen_bn_formula():
    bn = up.FrozenBatchNorm2d(2)
    bn.weight.copy_(torch.tensor([2.0, 0.5])); bn.bias.copy_(torch.tensor([1.0, -1.0]))
    bn.running_mean.copy_(torch.tensor([0.5, 0.0])); bn.running_var.copy_(torch.tensor([4.0, 1.0]))
    x = torch.tensor([1.5, 3.0]).view(1, 2, 1, 1)
    y = bn(x).view(-1)
    assert math.isclose(float(y[0]), (1.5 - 0.5) / math.sqrt(4 + 1e-5) * 2 + 1, rel_tol=1e-6)
    assert math.isclose(float(y[1]), 3.0 / math.sqrt(1 + 1e-5) * 0.5 - 1, rel_tol=1e-6)


def test_detector_postprocess_scales_clips_and_drops_empty():
    inst = up.Instances((100, 200))
    inst.pred_boxes = up.Boxes(torch.tensor([[10.0, 10.0, 50.0, 60.0], [190.0, 90.0, 260.0, 130.0], [-20.0, 5.0, -1.0, 9.0]]))
    inst.scores = torch.tensor([0.9, 0.8, 0.7])
    out = up.detector_postprocess(inst, 50, 100)
    assert len(out) == 2 and out.image_size == (50, 100)
    assert out.pred_boxes.tensor.tolist() == [[5.0, 5.0, 25.0, 30.0], [95.0, 45.0, 100.0, 50.0]]
