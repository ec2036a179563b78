"""The CPU oracle (oracle/meta_fcos_oracle.py) against the golden vectors produced by the REFERENCE's own modules
(oracle/make_golden.py ran them unmodified in the build container).  fp32 on CPU on both sides: agreement is expected
to the last bits; the tolerance only allows for a different BLAS summation order on another host."""
import pytest
import torch

from oracle.meta_fcos_oracle import MetaFCOSOracle
from sylph_few_shot_detection_b200 import weights as W
from tests.cases import cfg_for, load_golden, rel_err

TOL = 2e-5


@pytest.mark.parametrize("case", ["coco_2way_2shot", "lvis_1way_3shot", "coco_weight_layer_2way_3shot"])
def test_oracle_reproduces_reference_outputs(case):
    g = load_golden(case)
    cfg = cfg_for(g["config"], g.get("opts"))     # coco_weight_layer_*: CODE_GENERATOR.WEIGHT_LAYER ["", "", 1] (softmax shot weights)
    state = W.synthetic_state_dict(cfg, g["seed"])
    orc = MetaFCOSOracle(cfg, state)
    normed = []
    for c, shots in enumerate(g["support"]):
        code = orc.class_code([s["image"].float() for s in shots], torch.stack([s["box"] for s in shots]))
        assert code["cls_conv"].shape == (1, 256, 1, 1) and code["cls_bias"].shape == (1, 1, 1, 1)
        if "weight_layer" in case:   # the weight head is live: softmax weights far from 1 / K
            wts = orc.shot_weights(1, len(shots), torch.float32).reshape(-1)
            assert abs(float(wts.sum()) - 1.0) < 1e-6 and float(wts.max() - wts.min()) > 0.05
        assert rel_err(code["cls_conv"], g["raw_codes"][c]["cls_conv"]) < TOL
        assert rel_err(code["cls_bias"], g["raw_codes"][c]["cls_bias"]) < TOL
        w, b = orc.normalize_code(code["cls_conv"], code["cls_bias"])
        assert b.shape == (1,)
        assert rel_err(w, g["norm_codes"][c]["cls_conv"]) < TOL
        assert rel_err(b, g["norm_codes"][c]["cls_bias"]) < TOL
        normed.append({"support_set_target": c, "class_code": {"cls_conv": w, "cls_bias": b}})
    packed = orc.pack_codes(normed)
    assert packed["cls_conv"].shape == g["packed"]["cls_conv"].shape
    assert packed["cls_bias"].shape == g["packed"]["cls_bias"].shape
    dets, inter = orc.detect([q.float() for q in g["query"]], g["packed"], return_intermediate=True)
    for l in range(5):
        assert rel_err(inter["logits"][l], g["logits"][l]) < TOL
        assert rel_err(inter["reg"][l], g["reg"][l]) < TOL
        assert rel_err(inter["ctr"][l], g["ctr"][l]) < TOL
    for d, r in zip(dets, g["detections"]):
        assert d["scores"].numel() == r["scores"].numel() > 0
        assert torch.equal(d["classes"], r["classes"])
        assert torch.equal(d["levels"], r["levels"])
        assert torch.equal(d["locations"], r["locations"])
        assert rel_err(d["boxes"], r["boxes"]) < TOL
        assert rel_err(d["scores"], r["scores"]) < TOL


def test_roi_encoder_oracle_reproduces_reference_outputs():
    """ROIEncoder generator + CondConvBlock head against the reference's own modules."""
    from oracle.roi_encoder_oracle import ROIEncoderOracle, build_oracle
    g = load_golden("lvis_roienc_2way_3shot")
    cfg = cfg_for(g["config"], g["opts"])
    state = W.synthetic_state_dict(cfg, g["seed"])
    orc = build_oracle(cfg, state)
    assert isinstance(orc, ROIEncoderOracle)
    assert g["normalize_error"].startswith("TypeError")      # the reference cannot normalise ROIEncoder codes (quirk)
    codes = []
    for c, shots in enumerate(g["support"]):
        code = orc.class_code([s["image"].float() for s in shots], torch.stack([s["box"] for s in shots]))
        assert code["cls_conv"].shape == (1, 256, 1, 1) and code["cls_bias"].shape == (1,)
        assert rel_err(code["cls_conv"], g["raw_codes"][c]["cls_conv"]) < TOL
        assert rel_err(code["cls_bias"], g["raw_codes"][c]["cls_bias"]) < TOL
        codes.append({"support_set_target": c, "class_code": code})
    packed = orc.pack_codes(codes)
    assert torch.equal(packed["cls_conv"], g["packed"]["cls_conv"]) or rel_err(packed["cls_conv"], g["packed"]["cls_conv"]) < TOL
    # two classes in ONE call: bs = 2 -> the transformer attends across the CLASS axis (batch_first=False), so the
    # result differs from two bs = 1 calls; the restatement must still run and keep shapes (roi_encoder.py:176-199)
    both = orc.class_code([s["image"].float() for shots in g["support"] for s in shots],
                          torch.stack([s["box"] for shots in g["support"] for s in shots]))
    assert both["cls_conv"].shape == (2, 256, 1, 1) and both["cls_bias"].shape == (2,)
    assert rel_err(both["cls_conv"][:1], g["raw_codes"][0]["cls_conv"]) > 1e-3
    dets, inter = orc.detect([q.float() for q in g["query"]], g["packed"], return_intermediate=True)
    for l in range(5):
        assert rel_err(inter["logits"][l], g["logits"][l]) < TOL
        assert rel_err(inter["reg"][l], g["reg"][l]) < TOL
    for d, r in zip(dets, g["detections"]):
        assert d["scores"].numel() == r["scores"].numel() > 0
        assert torch.equal(d["classes"], r["classes"]) and torch.equal(d["levels"], r["levels"])
        assert torch.equal(d["locations"], r["locations"])
        assert rel_err(d["boxes"], r["boxes"]) < TOL and rel_err(d["scores"], r["scores"]) < TOL


def base_detector_state(cfg, seed):
    """The weights oracle/make_golden.py::base_detector_state gave the reference model (class-logits conv that fires)."""
    state = W.synthetic_state_dict(cfg, seed)
    k = "proposal_generator.fcos_head.cls_logits.weight"
    g = torch.Generator().manual_seed(4242 + seed)
    state[k] = torch.nn.functional.normalize(torch.randn(state[k].shape, generator=g), dim=1) * 5.0
    state["proposal_generator.fcos_head.cls_logits.bias"] = torch.randn(state[k].shape[0], generator=g) * 0.3 - 4.2
    return state


def test_oracle_reproduces_the_reference_base_detector():
    """`run_type=None` on the NON-episodic reference model (Meta-FCOS-pretrain.yaml; meta_one_stage_detector.py:298-323):
    the model's own `cls_logits` convolution is the classifier.  For the oracle these weights are simply the code rows."""
    g = load_golden("coco_base_detector")
    cfg = cfg_for(g["config"])
    assert not cfg.MODEL.META_LEARN.EPISODIC_LEARNING
    state = base_detector_state(cfg, g["seed"])
    assert not any(k.startswith("code_generator.") for k in state)
    orc = MetaFCOSOracle(cfg, state)
    codes = {"cls_conv": state["proposal_generator.fcos_head.cls_logits.weight"],
             "cls_bias": state["proposal_generator.fcos_head.cls_logits.bias"]}
    dets, inter = orc.detect([q.float() for q in g["query"]], codes, return_intermediate=True)
    for l in range(5):
        assert rel_err(inter["logits"][l], g["logits"][l]) < TOL and rel_err(inter["reg"][l], g["reg"][l]) < TOL
    for d, r in zip(dets, g["detections"]):
        assert d["scores"].numel() == r["scores"].numel() > 0
        assert torch.equal(d["classes"], r["classes"]) and torch.equal(d["locations"], r["locations"])
        assert rel_err(d["boxes"], r["boxes"]) < TOL and rel_err(d["scores"], r["scores"]) < TOL


def test_golden_vectors_cover_the_interesting_regimes():
    g = load_golden("coco_2way_2shot")
    # post-NMS top-k saturated on one image, not on the other; several FPN levels contribute
    counts = [int(d["scores"].numel()) for d in g["detections"]]
    assert max(counts) == 100 and min(counts) < 100
    assert len(set(int(v) for v in g["detections"][0]["levels"])) >= 2
    g = load_golden("lvis_1way_3shot")
    assert len(g["support"]) == 1 and len(g["support"][0]) == 3
