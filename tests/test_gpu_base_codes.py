"""GPU side of SURVEY.md 8f rows 1-2: the accumulation / reduction kernels through the C ABI are BIT-EXACT against the
golden vectors of the reference's reduce_class_code; the base-class path end to end against the oracle; the class-code
store + predictor against the direct plugin call."""
import copy

import numpy as np
import pytest
import torch

from tests.cases import load_golden, rel_err
from tests.test_gpu_cases import _images, _setup, _support_item

pytestmark = pytest.mark.gpu


def _same_code(a, b):
    return torch.equal(a["cls_conv"].cpu(), b["cls_conv"].cpu()) and torch.equal(a["cls_bias"].cpu().reshape(-1), b["cls_bias"].cpu().reshape(-1))


def test_accumulate_and_reduce_kernels_bit_exact_against_reference_golden():
    from sylph_few_shot_detection_b200.runner import reduce_class_code, replace_class_code
    cfg, state, model, orc = _setup()
    eng = model.engine
    g = load_golden("base_reduce")
    gathered = []
    for rk, ref in zip(g["chunks_per_rank"], g["per_rank"]):
        order = []
        for c in rk:
            if c["cid"] not in order:
                order.append(c["cid"])
        rows = torch.cat([torch.cat([c["code"]["cls_conv"].reshape(1, 256), c["code"]["cls_bias"].reshape(1, 1)], 1) for c in rk]).cuda()
        acc = torch.zeros((len(order), 257), device="cuda")
        # two calls: accumulation continues across calls exactly like the reference's running sums
        half = len(rk) // 2
        cls_idx = [order.index(c["cid"]) for c in rk]
        w = [float(c["len"]) / c["total_len"] for c in rk]
        eng.accumulate_codes(rows[:half], cls_idx[:half], w[:half], acc)
        eng.accumulate_codes(rows[half:], cls_idx[half:], w[half:], acc)
        assert [c["support_set_target"] for c in ref] == order
        for i, c in enumerate(ref):
            assert torch.equal(acc[i, :256].cpu(), c["class_code"]["cls_conv"].reshape(-1))
            assert torch.equal(acc[i, 256].cpu(), c["class_code"]["cls_bias"].reshape(()))
            gathered.append({"support_set_target": c["support_set_target"], "class_name": c["class_name"],
                             "class_code": {"cls_conv": acc[i, :256].reshape(1, 256, 1, 1), "cls_bias": acc[i, 256:].reshape(1, 1, 1, 1),
                                            "acc_weight": c["class_code"]["acc_weight"]}})
    reduced = reduce_class_code(gathered, eng)
    assert len(reduced) == len(g["reduced"])
    for a, b in zip(reduced, g["reduced"]):
        assert int(a["support_set_target"]) == int(b["support_set_target"]) and a["class_name"] == b["class_name"]
        assert _same_code(a["class_code"], b["class_code"]) and "acc_weight" not in a["class_code"]
    few = [dict(c, class_code={k: v.cuda() for k, v in c["class_code"].items()}) for c in g["few_shot"]]
    replaced = replace_class_code(few, reduced, torch.device("cuda"))
    for a, b in zip(replaced, g["replaced"]):
        assert _same_code(a["class_code"], b["class_code"]) and a["class_code"]["cls_conv"].is_cuda
    with pytest.raises(RuntimeError, match="class id"):
        eng.accumulate_codes(torch.zeros(1, 257), [99], [1.0], torch.zeros((2, 257), device="cuda"))


def test_base_class_path_end_to_end_against_oracle():
    """3 chunks of class 0 (4 + 4 + 2 boxes -> weights .4 .4 .2) and 1 chunk of class 1, through
    inference_on_support_set_base -> gather(reduce=True) -> replace_class_code -> normalise -> detect."""
    from oracle import base_codes_oracle as bo
    from sylph_few_shot_detection_b200.runner import (MetaFCOSRunner, format_class_codes_shared, inference_normalization,
                                                      inference_on_support_set, inference_on_support_set_base, replace_class_code)
    cfg, state, model, orc = _setup(seed=8)
    ims = _images(14, 160, 224, 3)
    rng = np.random.RandomState(2)

    def box():
        x0, y0 = rng.uniform(5, 60), rng.uniform(5, 40)
        return torch.tensor([x0, y0, x0 + rng.uniform(60, 150), y0 + rng.uniform(50, 110)], dtype=torch.float32)
    chunks, layout = [], [(0, 4, 10), (0, 4, 10), (1, 2, 2), (0, 2, 10)]
    used = 0
    for cid, ln, tot in layout:
        boxes = torch.stack([box() for _ in range(ln)])
        item = _support_item(ims[used:used + ln], boxes, cid)
        item.update({"len": ln, "total_len": tot})
        chunks.append((item, boxes, used))
        used += ln
    base = inference_on_support_set_base(model, [c[0] for c in chunks], chunks_per_batch=3)
    assert [c["support_set_target"] for c in base] == [0, 1]
    assert base[0]["class_code"]["acc_weight"] == pytest.approx(1.0) and base[1]["class_code"]["acc_weight"] == 1.0
    ref_chunks = [orc.class_code([im.float() for im in ims[u:u + b.shape[0]]], b) for _, b, u in chunks]
    ref = bo.accumulate_base_codes(ref_chunks, [l[0] for l in layout], [l[1] for l in layout], [l[2] for l in layout],
                                   [f"c{l[0]}" for l in layout])
    for a, b in zip(base, ref):
        assert rel_err(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"]) < 1e-3
        assert abs(float(a["class_code"]["cls_bias"]) - float(b["class_code"]["cls_bias"])) < 1e-3
    runner = MetaFCOSRunner()
    type(runner)._model = model
    base_codes = runner._gather_class_code(base, reduce=True)
    assert all("acc_weight" not in c["class_code"] for c in base_codes)
    few = inference_on_support_set(model, [_support_item(ims[12:14], torch.stack([box(), box()]), 2),
                                           _support_item(ims[0:2], torch.stack([box(), box()]), 0),
                                           _support_item(ims[2:4], torch.stack([box(), box()]), 1)])
    codes = replace_class_code(few, base_codes, model.device)
    assert torch.equal(codes[1]["class_code"]["cls_conv"], base_codes[0]["class_code"]["cls_conv"])   # class 0 replaced
    assert torch.equal(codes[0]["class_code"]["cls_conv"], few[0]["class_code"]["cls_conv"])           # class 2 kept
    packed = format_class_codes_shared(inference_normalization(model, codes), device=model.device)
    assert packed["cls_conv"].shape == (3, 256, 1, 1) and packed["cls_bias"].shape == (3,)
    out = model([{"image": ims[5], "height": 160, "width": 224}], class_code=packed, run_type="meta_learn_test_instance")
    assert out[0]["instances"].pred_boxes.tensor.shape[1] == 4


def test_code_store_and_predictor_match_direct_call(tmp_path):
    from sylph_few_shot_detection_b200.predictor import SylphPredictor, save_class_codes
    from sylph_few_shot_detection_b200.runner import format_class_codes_shared, inference_on_support_set
    cfg, state, model, orc = _setup(["INPUT.MIN_SIZE_TEST", 160, "INPUT.MAX_SIZE_TEST", 320], seed=4)
    ims = _images(5, 160, 224, 11)
    boxes = torch.tensor([[20.0, 30.0, 150.0, 140.0], [5.0, 5.0, 200.0, 150.0]])
    items = [_support_item(ims[0:2], boxes, 0), _support_item(ims[2:4], boxes, 1)]
    np.random.seed(0)
    res = inference_on_support_set(model, items)
    save_class_codes(res, str(tmp_path / "ds" / "0"))
    pred = SylphPredictor(cfg, state, str(tmp_path), {"all": ("ds", ["c0", "c1"])})
    direct = format_class_codes_shared(copy.deepcopy(res), device=model.device)
    assert torch.equal(pred.class_codes["all"]["cls_conv"], direct["cls_conv"])      # disk round trip is lossless
    assert torch.equal(pred.class_codes["all"]["cls_bias"], direct["cls_bias"])
    bgr = ims[4].permute(1, 2, 0).contiguous().numpy()                                # HxWx3 uint8, already 160 on the short side
    a = pred._call_few_shot(bgr, pred.class_codes["all"])["instances"]
    b = model([{"image": ims[4], "height": 160, "width": 224}], class_code=direct, run_type="meta_learn_test_instance")[0]["instances"]
    assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.scores, b.scores)
    # a 2x larger input is resized back to 160 on the short side and boxes come out in the ORIGINAL frame
    big = np.repeat(np.repeat(bgr, 2, axis=0), 2, axis=1)
    c = pred._call_few_shot(big, pred.class_codes["all"])["instances"]
    assert c.image_size == (320, 448)
    # incremental registration: the "user" split grows by one class per call
    np.random.seed(0)
    assert pred.register_class(items[0]) == 0 and pred.register_class(items[1]) == 1
    assert pred.class_codes["user"]["cls_conv"].shape == (2, 256, 1, 1)
    assert torch.equal(pred.class_codes["user"]["cls_conv"][0], direct["cls_conv"][0])
    with pytest.raises(ValueError, match="is missing"):
        SylphPredictor(cfg, state, str(tmp_path), {"all": ("ds", ["c0", "zebra"])})
