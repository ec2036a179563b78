"""More GPU parity / property tests through the plugin API and the C ABI: edge cases the reference's domain has
(no candidates, more candidates than PRE_NMS_TOPK, ragged image sizes, many classes, R-101, foreign NCHW features)
and size-independent properties at BASELINE.json's full 800x1333 size."""
import numpy as np
import pytest
import torch

from tests.cases import rel_err
from tests.parity import TOL, check_detections, instances_to_keyed

pytestmark = pytest.mark.gpu


def _setup(opts=None, seed=21, preset="COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", precision="exact"):
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import preset_cfg
    cfg = preset_cfg(preset, opts)
    state = W.synthetic_state_dict(cfg, seed)
    model = build_model(cfg, precision)
    model.load_state_dict(state)
    return cfg, state, model, MetaFCOSOracle(cfg, state)


def _images(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        base = torch.rand(3, h // 8 + 2, w // 8 + 2, generator=g) * 255.0
        img = torch.nn.functional.interpolate(base[None], size=(h, w), mode="bilinear", align_corners=False)[0]
        out.append((img + (torch.rand(3, h, w, generator=g) - 0.5) * 40.0).clamp(0, 255).round().to(torch.uint8))
    return out


def _support_item(images, boxes, cls):
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    recs = []
    for im, b in zip(images, boxes):
        inst = Instances(tuple(im.shape[-2:]))
        inst.gt_boxes = Boxes(b[None])
        inst.gt_classes = torch.tensor([cls])
        recs.append({"image": im, "instances": inst, "height": im.shape[-2], "width": im.shape[-1]})
    return {"support_set": recs, "support_set_target": torch.tensor(cls), "class_name": f"c{cls}"}


def _match(inst, ref, inter, image, cfg, box_tol_px=None, name=""):
    """Strict detection-set parity (tests/parity.py): matched boxes / scores within the bar, every key present on one
    side only explained by an oracle value inside the guard band of a decision threshold."""
    st = check_detections(instances_to_keyed(inst), ref, inter, image, cfg, box_tol_px=box_tol_px, name=name)
    return st["n_common"], st["n_diff"]


def test_no_candidates_gives_empty_instances_and_codes_still_match():
    """Reference initialisers leave every logit at the prior (sigmoid(-4.6) = 0.01 < 0.05): zero candidates."""
    cfg, state, model, orc = _setup()
    ims = _images(3, 160, 224, 1)
    box = torch.tensor([[20.0, 30.0, 120.0, 140.0], [5.0, 5.0, 200.0, 150.0]])
    code = model([_support_item(ims[:2], box, 0)], run_type="meta_learn_test_support")
    ref = orc.class_code([i.float() for i in ims[:2]], box)
    assert code["cls_conv"].shape == (1, 256, 1, 1) and code["cls_bias"].shape == (1, 1, 1, 1)
    assert rel_err(code["cls_conv"], ref["cls_conv"]) < 1e-3
    dead = {"cls_conv": torch.zeros(3, 256, 1, 1), "cls_bias": torch.full((3,), -4.59512)}
    out = model([{"image": ims[2], "height": 160, "width": 224}], class_code=dead, run_type="meta_learn_test_instance")
    inst = out[0]["instances"]
    assert len(inst) == 0 and inst.pred_boxes.tensor.shape == (0, 4) and inst.pred_classes.dtype == torch.int64
    assert inst.image_size == (160, 224)


def test_more_candidates_than_pre_nms_topk_and_many_classes():
    """Low threshold + small PRE_NMS_TOPK exercises the exact radix select; 20 classes exercise the 64-wide logits GEMM."""
    cfg, state, model, orc = _setup(["MODEL.FCOS.INFERENCE_TH_TEST", 0.002, "MODEL.FCOS.PRE_NMS_TOPK_TEST", 150,
                                     "MODEL.FCOS.POST_NMS_TOPK_TEST", 60], seed=4)
    g = torch.Generator().manual_seed(9)
    w = torch.nn.functional.normalize(torch.randn(20, 256, 1, 1, generator=g), dim=1) * 5.0
    codes = {"cls_conv": w, "cls_bias": torch.randn(20, generator=g) * 0.3 - 4.0}
    ims = _images(2, 192, 256, 3)
    ims[1] = ims[1][:, :170, :230].contiguous()  # ragged batch: padded to the common /32 size
    items = [{"image": im, "height": im.shape[-2], "width": im.shape[-1]} for im in ims]
    out = model(items, class_code=codes, run_type="meta_learn_test_instance")
    ref, inter = orc.detect([i.float() for i in ims], codes, return_intermediate=True)
    assert max(int(p["scores"].numel()) for p in inter["pre_nms"]) >= 150 * 2, "test must overflow the per-level top-k"
    for i, (o, r) in enumerate(zip(out, ref)):
        n_common, n_diff = _match(o["instances"], r, inter, i, cfg, name=f"topk overflow image {i}")
        assert n_common >= 40


def test_output_rescaling_to_requested_height_width():
    cfg, state, model, orc = _setup(seed=4)
    g = torch.Generator().manual_seed(2)
    codes = {"cls_conv": torch.nn.functional.normalize(torch.randn(2, 256, 1, 1, generator=g), dim=1) * 6.0,
             "cls_bias": torch.tensor([-3.5, -3.8])}
    im = _images(1, 160, 256, 5)[0]
    out = model([{"image": im, "height": 320, "width": 512}], class_code=codes, run_type="meta_learn_test_instance")
    ref, inter = orc.detect([im.float()], codes, out_sizes=[(320, 512)], return_intermediate=True)
    assert out[0]["instances"].image_size == (320, 512)
    _match(out[0]["instances"], ref[0], inter, 0, cfg, box_tol_px=TOL * 512)
    b = out[0]["instances"].pred_boxes.tensor
    assert float(b[:, 0::2].max()) <= 512 and float(b[:, 1::2].max()) <= 320 and float(b.min()) >= 0


def test_resnet101_backbone_features():
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    cfg, state, model, orc = _setup(["MODEL.RESNETS.DEPTH", 101], seed=2)
    ims = _images(2, 128, 160, 8)
    model.engine.extract_features(SLOT_SUPPORT, [i.cuda() for i in ims])
    ref = orc.features(orc.preprocess([i.float() for i in ims]).tensor)
    for l in range(5):
        assert rel_err(model.engine.export_features(SLOT_SUPPORT, l), ref[l]) < TOL, l


def test_code_generator_plugin_with_foreign_nchw_features():
    """`CodeGenerator.forward(features, target_instances)` -- the plugin boundary -- fed with NCHW features computed
    elsewhere (here: the CPU oracle's backbone), like tests/code_generator_roi_encoder_test.py does with synthetic
    features in the reference."""
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    cfg, state, model, orc = _setup(seed=6, preset="LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml")
    ims = _images(3, 256, 320, 12)
    feats = orc.features(orc.preprocess([i.float() for i in ims]).tensor)
    boxes = torch.tensor([[30.0, 40.0, 90.0, 120.0], [10.0, 10.0, 300.0, 250.0], [100.0, 60.0, 220.0, 200.0]])
    insts = []
    for b in boxes:
        inst = Instances((256, 320))
        inst.gt_boxes = Boxes(b[None])
        insts.append(inst)
    out = model.code_generator([f.cuda() for f in feats], insts)
    roi, _ = orc.roi_features(feats, boxes)
    w, b = orc.per_shot_codes(roi)
    assert out["cls_conv"].shape == (1, 256, 1, 1) and out["cls_bias"].shape == (1, 1, 1, 1)
    assert rel_err(out["cls_conv"].reshape(-1), w.mean(0).reshape(-1)) < 1e-3
    assert abs(float(out["cls_bias"]) - float(b.mean())) < 1e-3
    normed = model.normalize_class_code([{"support_set_target": 0, "class_code": dict(out)}])
    wn, bn = orc.normalize_code(w.mean(0, keepdim=True), b.mean(0, keepdim=True))
    assert normed[0]["class_code"]["cls_bias"].shape == (1,)
    assert rel_err(normed[0]["class_code"]["cls_conv"], wn) < 1e-3 and rel_err(normed[0]["class_code"]["cls_bias"], bn) < 1e-3


def test_backbone_plugin_on_a_normalised_batch_feeds_the_code_generator_plugin():
    """`build_fcos_resnet_fpn_backbone(cfg)(images.tensor)` as the reference calls it (meta_one_stage_detector.py:174-182):
    a normalised, zero-padded (N, 3, H, W) batch in, {p3..p7} NCHW fp32 out; the result goes straight into the
    reference-shaped `CodeGenerator.forward(features, target_instances)`."""
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    cfg, state, model, orc = _setup(seed=6)
    ims = _images(2, 150, 200, 14)                       # pads to 160 x 224
    il = orc.preprocess([i.float() for i in ims])        # (x - mean) / std, zero padding: what the reference hands over
    ref = orc.features(il.tensor)
    out = model.backbone(il.tensor.cuda())
    assert list(out) == ["p3", "p4", "p5", "p6", "p7"]
    for l, name in enumerate(out):
        assert out[name].shape == ref[l].shape and out[name].dtype == torch.float32
        assert rel_err(out[name], ref[l]) < TOL, name
    # identical to the fused raw-image entry (the normalisation of uint8 pixels is exact in both)
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    model.engine.extract_features(SLOT_QUERY, [i.cuda() for i in ims])
    for l, name in enumerate(out):
        assert rel_err(model.engine.export_features(SLOT_QUERY, l), out[name]) < 1e-5, name
    boxes = torch.tensor([[30.0, 40.0, 90.0, 120.0], [10.0, 10.0, 190.0, 140.0]])
    insts = []
    for b in boxes:
        inst = Instances((150, 200))
        inst.gt_boxes = Boxes(b[None])
        insts.append(inst)
    code = model.code_generator([out[f] for f in cfg.MODEL.FCOS.IN_FEATURES], insts)
    want = orc.class_code([i.float() for i in ims], boxes)
    assert rel_err(code["cls_conv"], want["cls_conv"]) < TOL
    assert abs(float(code["cls_bias"]) - float(want["cls_bias"])) < TOL


def test_proposal_generator_on_foreign_features_returns_unscaled_proposals():
    """`MetaFCOS.forward(images, features, support_set_per_class_code=...)` with NCHW features of a padded
    batch whose images are SMALLER than the padded size: proposals come back un-scaled (the reference's proposal
    generator never rescales; fcos_outputs.py:986-1006), clipped to the image."""
    from oracle import upstream as up
    cfg, state, model, orc = _setup(seed=4)
    g = torch.Generator().manual_seed(2)
    codes = {"cls_conv": torch.nn.functional.normalize(torch.randn(2, 256, 1, 1, generator=g), dim=1) * 6.0,
             "cls_bias": torch.tensor([-3.5, -3.8])}
    ims = [_images(1, 150, 199, 5)[0], _images(1, 140, 210, 6)[0]]        # batch pads to 160 x 224
    il = orc.preprocess([i.float() for i in ims])
    feats = orc.features(il.tensor)
    features = {f"p{3 + l}": f.cuda() for l, f in enumerate(feats)}
    props, losses = model.proposal_generator(up.ImageList(il.tensor, il.image_sizes), features, support_set_per_class_code=codes)
    assert losses == {} and len(props) == 2
    ref, inter = orc.detect([i.float() for i in ims], codes, return_intermediate=True)   # out size == image size: scale 1
    for i, (p, r) in enumerate(zip(props, ref)):
        assert p.image_size == tuple(il.image_sizes[i])
        _match(p, r, inter, i, cfg, name=f"foreign features image {i}")


def test_base_detector_inference_matches_the_reference_golden():
    """`model(batched_inputs)` with run_type=None on the NON-episodic model (Meta-FCOS-pretrain.yaml): the reference's
    "normal base detector inference" (meta_one_stage_detector.py:298-323, 435-441) against the golden its own model produced;
    an episodic model refuses the call with the reference's message."""
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    from tests.cases import cfg_for, load_golden
    from tests.test_oracle import base_detector_state
    g = load_golden("coco_base_detector")
    cfg = cfg_for(g["config"])
    state = base_detector_state(cfg, g["seed"])
    model = build_model(cfg)
    assert model.code_generator is None
    model.load_state_dict(state)
    batched = [{"image": q, "height": q.shape[-2], "width": q.shape[-1]} for q in g["query"]]
    out = model(batched)
    assert len(out) == 2 and set(out[0]) == {"instances"}
    orc = MetaFCOSOracle(cfg, state)
    codes = {"cls_conv": state["proposal_generator.fcos_head.cls_logits.weight"], "cls_bias": state["proposal_generator.fcos_head.cls_logits.bias"]}
    ref, inter = orc.detect([q.float() for q in g["query"]], codes, return_intermediate=True)
    for l in range(5):
        assert rel_err(model.engine.export_head_output(0, l, SLOT_QUERY, 60), g["logits"][l]) < TOL, l
    for i, (o, r) in enumerate(zip(out, ref)):
        assert r["scores"].numel() == g["detections"][i]["scores"].numel()
        _match(o["instances"], r, inter, i, cfg, name=f"base detector image {i}")
    with pytest.raises(RuntimeError, match="no code generator"):
        model.engine.normalize_codes(torch.zeros(1, 257))
    _, _, episodic, _ = _setup(seed=4)
    with pytest.raises(NotImplementedError, match="Episodic learning inferrence"):
        episodic(batched)


def test_batched_class_codes_equal_per_class_calls_and_levels_are_exact():
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    from oracle import upstream as up
    cfg, state, model, orc = _setup(seed=3)
    ims = _images(4, 480, 640, 17)
    boxes = torch.tensor([[10.0, 10.0, 100.0, 90.0], [0.0, 0.0, 639.0, 479.0], [100.0, 100.0, 420.0, 400.0], [300.0, 200.0, 332.0, 230.0]])
    items = [_support_item(ims[:2], boxes[:2], 0), _support_item(ims[2:], boxes[2:], 1)]
    np.random.seed(0)
    batched = model.forward_class_codes_batched(items)
    single = [model([it], run_type="meta_learn_test_support") for it in items]
    for a, b in zip(batched, single):
        assert torch.equal(a["cls_conv"], b["cls_conv"]) and torch.equal(a["cls_bias"], b["cls_bias"])
    model.engine.extract_features(SLOT_SUPPORT, [i.cuda() for i in ims])
    _, levels = model.engine.generate_codes(SLOT_SUPPORT, boxes, [0, 1, 2, 3], [0, 2, 4], want_levels=True)
    ref = up.assign_boxes_to_levels([up.Boxes(b[None]) for b in boxes], 3, 7, 224, 4)
    assert levels.dtype == torch.int64 and torch.equal(levels.cpu(), ref)
    assert len(set(ref.tolist())) >= 3  # boxes were chosen to land on several FPN levels


def test_batched_class_codes_with_different_image_sizes_per_class_and_chunked_passes():
    """the reference pads each class call to its own maximum, so classes of different (padded) sizes must not
    share a trunk batch; and a long class list is walked in bounded passes.  Batched == per-class, bit for bit."""
    cfg, state, model, orc = _setup(seed=3)
    a, b, c = _images(2, 160, 224, 31), _images(2, 128, 200, 32), _images(2, 150, 210, 33)    # pad to 160x224, 128x224, 160x224
    box = torch.tensor([[10.0, 10.0, 100.0, 90.0], [20.0, 30.0, 180.0, 120.0]])
    items = [_support_item(a, box, 0), _support_item(b, box, 1), _support_item(c, box, 2)]
    sizes = [model.class_padded_size(it) for it in items]
    assert sizes == [(160, 224), (128, 224), (160, 224)]
    np.random.seed(0)
    single = [model([it], run_type="meta_learn_test_support") for it in items]
    np.random.seed(0)
    batched = model.forward_class_codes_batched(items)
    for x, y in zip(batched, single):
        assert torch.equal(x["cls_conv"], y["cls_conv"]) and torch.equal(x["cls_bias"], y["cls_bias"])
    ref = orc.class_code([i.float() for i in b], box)        # the odd-sized class against the oracle's own padding
    assert rel_err(batched[1]["cls_conv"], ref["cls_conv"]) < TOL
    # a budget of ~1 full-size image per pass: every class becomes its own pass, results unchanged
    model.SUPPORT_PASS_IMAGE_BUDGET = 0.05
    try:
        np.random.seed(0)
        chunked = model.forward_class_codes_batched(items)
    finally:
        del model.SUPPORT_PASS_IMAGE_BUDGET
    for x, y in zip(chunked, single):
        assert torch.equal(x["cls_conv"], y["cls_conv"]) and torch.equal(x["cls_bias"], y["cls_bias"])


def test_full_size_episode_properties():
    """BASELINE configs[1] size (800x1333): no oracle run (minutes on CPU); size-independent properties instead --
    run-to-run determinism, descending scores, boxes inside the image, post-NMS count bound, finite class codes."""
    from sylph_few_shot_detection_b200.runner import run_episode
    cfg, state, model, orc = _setup(seed=0)
    g = torch.Generator().manual_seed(77)
    ims = [torch.randint(0, 256, (3, 800, 1333), generator=g, dtype=torch.uint8) for _ in range(6)]
    boxes = torch.tensor([[100.0, 100.0, 400.0, 500.0], [600.0, 200.0, 1300.0, 780.0], [50.0, 40.0, 120.0, 130.0], [0.0, 0.0, 1333.0, 800.0]])
    support = [_support_item(ims[:2], boxes[:2], 0), _support_item(ims[2:4], boxes[2:], 1)]
    query = [{"image": im, "height": 800, "width": 1333} for im in ims[4:]]
    a = run_episode(model, support, query)
    b = run_episode(model, support, query)
    for ra, rb in zip(a, b):
        ia, ib = ra["instances"], rb["instances"]
        assert torch.equal(ia.pred_boxes.tensor, ib.pred_boxes.tensor) and torch.equal(ia.scores, ib.scores)
        assert torch.equal(ia.pred_classes, ib.pred_classes)
        s = ia.scores.cpu()
        assert len(ia) <= 100 + 5 and bool((s[:-1] >= s[1:]).all()) and bool(torch.isfinite(s).all())
        bx = ia.pred_boxes.tensor.cpu()
        assert float(bx.min()) >= 0 and float(bx[:, 0::2].max()) <= 1333 and float(bx[:, 1::2].max()) <= 800
        assert bool(((bx[:, 2] > bx[:, 0]) & (bx[:, 3] > bx[:, 1])).all())
        assert int(ia.pred_classes.max()) <= 1 and int(ia.fpn_levels.max()) <= 4


def test_many_classes_use_the_wide_logits_gemm_and_class_sweep_codes():
    """Config-4/5 regimes at reduced size: 300 class codes (two 256-wide N tiles of the conditional classifier) and a
    class sweep of 40 classes x 3 shots generated in ONE launch sequence from a small pool of support images."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    cfg, state, model, orc = _setup(seed=8)
    g = torch.Generator().manual_seed(4)
    n_cls = 300
    codes = {"cls_conv": torch.nn.functional.normalize(torch.randn(n_cls, 256, 1, 1, generator=g), dim=1) * 4.0,
             "cls_bias": torch.randn(n_cls, generator=g) * 0.2 - 4.5}
    im = _images(1, 128, 192, 6)[0]
    out = model([{"image": im, "height": 128, "width": 192}], class_code=codes, run_type="meta_learn_test_instance")
    ref, inter = orc.detect([im.float()], codes, return_intermediate=True)
    for l in range(5):
        got = model.engine.export_head_output(0, l, SLOT_QUERY, n_cls)
        assert got.shape == inter["logits"][l].shape
        assert rel_err(got, inter["logits"][l]) < TOL, l
    _match(out[0]["instances"], ref[0], inter, 0, cfg, name="300 classes")
    assert int(out[0]["instances"].pred_classes.max()) > 255  # classes of the second N tile are reachable

    pool = _images(4, 256, 320, 40)
    n_classes, shots = 40, 3
    boxes, roi_image, offsets = [], [], [0]
    for c in range(n_classes):
        for s in range(shots):
            side = 24.0 + 5.0 * ((c * shots + s) % 40)
            x0, y0 = float((7 * c + 3 * s) % 60), float((5 * c + 11 * s) % 40)
            boxes.append([x0, y0, min(x0 + side * 1.2, 319.0), min(y0 + side, 255.0)])
            roi_image.append((c + s) % 4)
        offsets.append(len(boxes))
    boxes = torch.tensor(boxes)
    model.engine.extract_features(SLOT_SUPPORT, [p.cuda() for p in pool])
    raw = model.engine.generate_codes(SLOT_SUPPORT, boxes, roi_image, offsets)
    feats = orc.features(orc.preprocess([p.float() for p in pool]).tensor)
    for c in (0, 17, 39):
        idx = list(range(offsets[c], offsets[c + 1]))
        sub = [f[[roi_image[i] for i in idx]] for f in feats]
        roi, _ = orc.roi_features(sub, boxes[idx])
        w, b = orc.per_shot_codes(roi)
        assert rel_err(raw[c, :256], w.mean(0).reshape(-1)) < 1e-3
        assert abs(float(raw[c, 256]) - float(b.mean())) < 1e-3


@pytest.mark.parametrize("sizes", [[(64, 64)], [(96, 160), (80, 150)], [(160, 96)], [(257, 131), (200, 97), (31, 33)],
                                   [(352, 480)], [(32, 512)], [(416, 64), (400, 40)]])
def test_geometry_sweep_features_and_head_outputs(sizes):
    """Odd, tiny, ragged and extreme-aspect images: every plane geometry (padding to /32, p6/p7 of 1-3 pixels, planes
    narrower than a tile, batches of different sizes) against the CPU oracle."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    cfg, state, model, orc = _setup(seed=5)
    ims = [_images(1, h, w, 100 + i)[0] for i, (h, w) in enumerate(sizes)]
    g = torch.Generator().manual_seed(3)
    codes = {"cls_conv": torch.nn.functional.normalize(torch.randn(3, 256, 1, 1, generator=g), dim=1) * 5.0,
             "cls_bias": torch.tensor([-3.0, -3.5, -4.0])}
    items = [{"image": im, "height": im.shape[-2], "width": im.shape[-1]} for im in ims]
    out = model(items, class_code=codes, run_type="meta_learn_test_instance")
    ref, inter = orc.detect([i.float() for i in ims], codes, return_intermediate=True)
    for l in range(5):
        got = model.engine.export_features(SLOT_QUERY, l)
        assert got.shape == inter["features"][l].shape, (l, got.shape, inter["features"][l].shape)
        assert rel_err(got, inter["features"][l]) < TOL, ("features", l)
        assert rel_err(model.engine.export_head_output(0, l, SLOT_QUERY, 3), inter["logits"][l]) < TOL, ("logits", l)
        assert rel_err(model.engine.export_head_output(1, l, SLOT_QUERY, 3), inter["reg"][l]) < TOL, ("reg", l)
        assert rel_err(model.engine.export_head_output(2, l, SLOT_QUERY, 3), inter["ctr"][l]) < TOL, ("ctr", l)
    for i, (o, r) in enumerate(zip(out, ref)):
        _match(o["instances"], r, inter, i, cfg, name=f"geometry {sizes} image {i}")


def test_merged_trunk_pass_is_bit_identical_to_separate_passes():
    """sylph_extract_features_multi: support + query batches through ONE trunk pass give exactly the pyramids of two
    separate calls (images are independent in the trunk); batches that pad to different sizes fall back to two passes
    (each reference call pads to its own batch maximum), so results are identical there as well."""
    from sylph_few_shot_detection_b200.runner import run_episode
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    cfg, state, model, orc = _setup(seed=6)
    eng = model.engine
    sup = [im.cuda() for im in _images(3, 160, 224, 5)]
    for qry in ([im.cuda() for im in _images(2, 150, 200, 6)],       # pads to 160 x 224 as well -> shared trunk batch
                [im.cuda() for im in _images(2, 96, 128, 7)]):       # pads to 96 x 128 -> fallback
        eng.extract_features(SLOT_SUPPORT, sup)
        eng.extract_features(SLOT_QUERY, qry)
        want = [[eng.export_features(s, l).clone() for l in range(5)] for s in (SLOT_SUPPORT, SLOT_QUERY)]
        eng.extract_features_multi([(SLOT_SUPPORT, sup), (SLOT_QUERY, qry)])
        for si, s in enumerate((SLOT_SUPPORT, SLOT_QUERY)):
            assert eng.feature_shape(s)[0] == (3 if s == SLOT_SUPPORT else 2)
            for l in range(5):
                assert torch.equal(eng.export_features(s, l), want[si][l])
    with pytest.raises(RuntimeError, match="listed twice"):
        eng.extract_features_multi([(SLOT_SUPPORT, sup), (SLOT_SUPPORT, sup)])
    # run_episode takes the merged path for device-resident inputs and must equal the host-resident (two-pass) run
    ims = _images(5, 160, 224, 9)
    boxes = torch.tensor([[20.0, 30.0, 120.0, 140.0], [5.0, 5.0, 200.0, 150.0]])
    np.random.seed(0)
    host = run_episode(model, [_support_item(ims[0:2], boxes, 0), _support_item(ims[2:4], boxes, 1)],
                       [{"image": ims[4], "height": 160, "width": 224}])
    np.random.seed(0)
    dev = run_episode(model, [_support_item([i.cuda() for i in ims[0:2]], boxes, 0), _support_item([i.cuda() for i in ims[2:4]], boxes, 1)],
                      [{"image": ims[4].cuda(), "height": 160, "width": 224}])
    assert torch.equal(host[0]["instances"].pred_boxes.tensor, dev[0]["instances"].pred_boxes.tensor)
    assert torch.equal(host[0]["instances"].scores, dev[0]["instances"].scores)


def test_cuda_graph_episode_replays_bit_identically():
    """runner.EpisodeGraph: the whole episode captured into one CUDA graph (PDL edges included); replays with new
    images / boxes written into the static buffers equal the eager calls bit for bit."""
    from sylph_few_shot_detection_b200.runner import EpisodeGraph
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    cfg, state, model, orc = _setup(seed=4)
    eng = model.engine
    g = EpisodeGraph(model, n_way=2, n_shot=2, n_query=2, image_hw=(160, 224))
    for rep, seed in enumerate((3, 8)):
        ims = [im.cuda() for im in _images(6, 160, 224, seed)]
        boxes = torch.tensor([[20.0, 30.0, 120.0, 140.0], [5.0, 5.0, 200.0, 150.0], [40.0, 20.0, 180.0, 100.0],
                              [60.0 + 5 * rep, 50.0, 140.0, 150.0]])
        for dst, src in zip(g.support, ims[:4]):
            dst.copy_(src)
        for dst, src in zip(g.query, ims[4:]):
            dst.copy_(src)
        g.boxes.copy_(boxes)
        dets, counts = g.replay()
        dets, counts = dets.clone(), counts.clone()
        eng.extract_features(SLOT_SUPPORT, ims[:4])
        raw = eng.generate_codes(SLOT_SUPPORT, boxes, [0, 1, 2, 3], [0, 2, 4])
        codes = eng.normalize_codes(raw)
        assert torch.equal(codes, g.codes)
        eng.extract_features(SLOT_QUERY, ims[4:])
        d2, c2 = eng.detect(SLOT_QUERY, codes)
        assert torch.equal(counts, c2) and torch.equal(dets, d2)
        assert int(counts.sum()) > 0


def test_uint8_images_give_bit_identical_features_to_fp32_images():
    """The uint8 prep kernel (table lookup + shared-memory staging with aligned word loads) against the fp32 prep
    kernel on the same pixel values: ragged sizes in one batch, odd widths (row starts at every byte alignment),
    a width beyond one 512-pixel segment, and images that are unaligned VIEWS into a larger byte buffer (first and
    last word straddle the tensor ends)."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    cfg, state, model, _ = _setup()
    eng = model.engine
    g = torch.Generator().manual_seed(77)
    sizes = [(97, 131), (64, 1333), (130, 257), (33, 70)]
    ims = []
    for k, (h, w) in enumerate(sizes):
        raw = torch.randint(0, 256, (3 * h * w + 16,), generator=g, dtype=torch.uint8).cuda()
        ims.append(raw[k + 1:k + 1 + 3 * h * w].view(3, h, w))      # storage offset 1, 2, 3, 4 bytes: unaligned views
    eng.extract_features(SLOT_QUERY, ims)
    a = [eng.export_features(SLOT_QUERY, l).clone() for l in range(5)]
    eng.extract_features(SLOT_QUERY, [im.float() for im in ims])
    b = [eng.export_features(SLOT_QUERY, l) for l in range(5)]
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_async_episode_pipeline_returns_the_same_detections():
    """EpisodePipeline.run_async / EpisodeFuture.result (no host synchronisation between episodes, detections copied
    to pinned memory) against the synchronous EpisodePipeline.run on the same host-resident episodes."""
    from sylph_few_shot_detection_b200.runner import EpisodePipeline
    cfg, state, model, _ = _setup()
    ims = _images(7, 160, 224, 5)
    boxes = torch.tensor([[20.0, 30.0, 150.0, 120.0], [5.0, 10.0, 200.0, 150.0]])
    episodes = []
    for e in range(3):
        support = [_support_item(ims[e:e + 2], boxes, 0), _support_item(ims[e + 2:e + 4], boxes, 1)]
        query = [{"image": ims[(e + 4) % 7], "height": 160, "width": 224}, {"image": ims[(e + 5) % 7], "height": 80, "width": 112}]
        episodes.append((support, query))
    pipe = EpisodePipeline(model)
    want = [pipe.run(pipe.submit(s, q)) for s, q in episodes]
    futures = [pipe.run_async(pipe.submit(s, q)) for s, q in episodes]      # all three enqueued before any result is read
    got = [f.result() for f in futures]
    for w_ep, g_ep in zip(want, got):
        assert len(w_ep) == len(g_ep) == 2
        for w, gi in zip(w_ep, g_ep):
            wi, gg = w["instances"], gi["instances"]
            assert not gg.scores.is_cuda and gg.image_size == wi.image_size
            assert torch.equal(wi.pred_boxes.tensor.cpu(), gg.pred_boxes.tensor)
            assert torch.equal(wi.scores.cpu(), gg.scores)
            assert torch.equal(wi.pred_classes.cpu(), gg.pred_classes)
            assert torch.equal(wi.fpn_levels.cpu(), gg.fpn_levels)


def test_codes_on_a_side_stream_give_identical_detections():
    """sylph_detect_after: code generation on the engine's side stream, towers on the current stream, joined by an event
    right before the code-conditioned classifier -- against the single-stream order, bit for bit."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    cfg, state, model, _ = _setup()
    eng = model.engine
    ims = [im.cuda() for im in _images(6, 192, 256, 9)]
    boxes = torch.tensor([[20.0, 30.0, 150.0, 120.0], [5.0, 10.0, 200.0, 150.0], [40.0, 40.0, 120.0, 160.0], [60.0, 20.0, 250.0, 100.0]])
    eng.extract_features_multi([(SLOT_SUPPORT, ims[:4]), (SLOT_QUERY, ims[4:])])
    (d0, c0), codes0 = eng.generate_and_detect(SLOT_SUPPORT, SLOT_QUERY, boxes, [0, 1, 2, 3], [0, 2, 4], overlap=False)
    d0, c0, codes0 = d0.clone(), c0.clone(), codes0.clone()
    for _ in range(3):
        (d1, c1), codes1 = eng.generate_and_detect(SLOT_SUPPORT, SLOT_QUERY, boxes, [0, 1, 2, 3], [0, 2, 4], overlap=True)
        torch.cuda.synchronize()
        assert torch.equal(codes0, codes1) and torch.equal(c0, c1) and torch.equal(d0, d1)
    assert int(c0.sum()) > 0


def test_non_integer_float_images_take_the_hi_lo_stem_path():
    """Exact mode re-centres pixels on round(pixel_mean): uint8 images need one stem pass (a_lo == 0); float images with
    fractional pixel values carry a lo half and run both passes -- features against the CPU oracle on the same float values,
    ragged sizes in one batch (padding value rn16((mean - round(mean)) / 256) outside each image and in the border ring)."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    cfg, state, model, orc = _setup(seed=9)
    g = torch.Generator().manual_seed(31)
    ims = []
    for i, (h, w) in enumerate([(120, 200), (96, 131), (160, 77)]):
        ims.append((_images(1, h, w, 300 + i)[0].float() + torch.rand(3, h, w, generator=g) * 0.98 - 0.49).clamp(0, 255))
    model.engine.extract_features(SLOT_QUERY, [im.cuda() for im in ims])
    codes = {"cls_conv": torch.zeros(1, 256, 1, 1), "cls_bias": torch.tensor([-5.0])}
    _, inter = orc.detect(ims, codes, return_intermediate=True)
    feats = inter["features"]
    for l in range(5):
        got = model.engine.export_features(SLOT_QUERY, l)
        assert got.shape == feats[l].shape
        assert rel_err(got, feats[l]) < TOL, ("features", l, rel_err(got, feats[l]))


@pytest.mark.parametrize("switch", ["SYLPH_NM", "SYLPH_QS", "SYLPH_PAIR1X1"])
def test_exact_mode_kernel_switches_give_the_same_features(switch, monkeypatch):
    """The exact-mode schedule switches (N-merged 3x3 / stem, quad stages, CTA-pair split 1x1 kernel) change which kernels run
    and the order of fp32 additions, never the arithmetic: features with a switch off agree with the default build to 2e-5
    (and both meet the 1e-3 bar against the oracle elsewhere in this suite)."""
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    ims = [im.cuda() for im in _images(2, 224, 320, 41)]
    _, _, model, _ = _setup(seed=4)
    model.engine.extract_features(SLOT_QUERY, ims)
    ref = [model.engine.export_features(SLOT_QUERY, l).clone() for l in range(5)]
    del model
    monkeypatch.setenv(switch, "0")
    _, _, model2, _ = _setup(seed=4)
    model2.engine.extract_features(SLOT_QUERY, ims)
    for l in range(5):
        got = model2.engine.export_features(SLOT_QUERY, l)
        assert rel_err(got, ref[l]) < 2e-5, (switch, l, rel_err(got, ref[l]))


@pytest.mark.parametrize("switch", ["SYLPH_ROI_PACKED", "SYLPH_CLS_POOLED"])
def test_code_generator_schedule_switches_give_the_same_codes(switch, monkeypatch):
    """Packed ROI planes (81 rows per ROI, per-ROI GroupNorm kernel) against one 128-row tile per ROI, and the pooled cls
    GEMM against the per-pixel cls convolution: the same arithmetic in another order -- raw class codes agree to 2e-5, and the
    default build meets the 1e-3 bar against the oracle in the parity tests."""
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    ims = [im.cuda() for im in _images(3, 224, 320, 51)]
    boxes = torch.tensor([[20.0, 30.0, 200.0, 180.0], [5.0, 5.0, 310.0, 215.0], [100.0, 60.0, 160.0, 120.0],
                          [40.0, 20.0, 120.0, 200.0], [10.0, 100.0, 300.0, 140.0], [150.0, 10.0, 250.0, 90.0], [60.0, 60.0, 90.0, 95.0]])
    roi_image, offsets = [0, 1, 2, 0, 1, 2, 0], [0, 3, 7]

    def codes():
        _, _, model, _ = _setup(seed=6)
        model.engine.extract_features(SLOT_SUPPORT, ims)
        return model.engine.generate_codes(SLOT_SUPPORT, boxes, roi_image, offsets).clone()
    ref = codes()
    monkeypatch.setenv(switch, "0")
    got = codes()
    assert rel_err(got[:, :256], ref[:, :256]) < 2e-5, (switch, rel_err(got[:, :256], ref[:, :256]))
    assert float((got[:, 256] - ref[:, 256]).abs().max()) < 2e-5
