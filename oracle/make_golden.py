"""ORACLE tooling (build container only): run the REFERENCE's own modules (unmodified, from /root/reference, on the
stand-in dependencies in oracle/shims) on small seeded episodes and store inputs + outputs as golden vectors.

    python -m oracle.make_golden            # writes tests/golden/*.pt and prints the oracle-vs-reference deltas

The weights are `sylph_few_shot_detection_b200.weights.synthetic_state_dict(cfg, seed)` loaded into the reference
model, so a test on the GPU box (no /root/reference there) can rebuild the identical model from (config, seed).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import reference_loader, upstream as up  # noqa: E402
from oracle.meta_fcos_oracle import MetaFCOSOracle  # noqa: E402
from oracle.roi_encoder_oracle import build_oracle  # noqa: E402
from sylph_few_shot_detection_b200 import weights as W  # noqa: E402
from sylph_few_shot_detection_b200.config import load_cfg  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
CONFIGS = {
    "coco": "COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml",
    "lvis": "LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml",
    "lvis_roienc": "LVISv1-Detection/Meta-FCOS/Meta-FCOS-ROI-Encoder-finetune.yaml",
}
# extra config overrides of a case (also applied by the tests that rebuild the cfg from the preset)
CASE_OPTS = {"lvis_roienc_2way_3shot": ["MODEL.META_LEARN.EVAL_SHOT", 3],
             # per-shot weight head (commented out in the shipped LVIS configs: "# WEIGHT_LAYER: [\"\", \"\", 1]")
             "coco_weight_layer_2way_3shot": ["MODEL.META_LEARN.CODE_GENERATOR.WEIGHT_LAYER", ["", "", 1]]}
# vendored copies of the two YAML trees are NOT kept; tests rebuild the cfg from these overrides on top of defaults
CASES = {
    # name: (config, seed, classes, shots, support (H, W) list, query (H, W) list)
    "coco_2way_2shot": ("coco", 3, 2, 2, [(256, 320), (240, 300)], [(256, 320), (200, 288)]),
    "lvis_1way_3shot": ("lvis", 5, 1, 3, [(224, 256), (256, 224), (200, 240)], [(224, 288)]),
    "lvis_roienc_2way_3shot": ("lvis_roienc", 9, 2, 3, [(224, 256), (256, 224), (200, 240)], [(224, 288)]),
    "coco_weight_layer_2way_3shot": ("coco", 13, 2, 3, [(224, 256), (256, 224), (200, 240)], [(224, 288)]),
}


def synth_image(g: torch.Generator, h: int, w: int) -> torch.Tensor:
    """uint8-valued fp32 BGR image, smooth + noise so that features are not white noise."""
    base = torch.rand(3, h // 8 + 2, w // 8 + 2, generator=g) * 255.0
    img = torch.nn.functional.interpolate(base[None], size=(h, w), mode="bilinear", align_corners=False)[0]
    img = img + (torch.rand(3, h, w, generator=g) - 0.5) * 40.0
    return img.clamp(0, 255).round()


def synth_box(g: torch.Generator, h: int, w: int, big: bool) -> torch.Tensor:
    side = (0.75 + 0.2 * torch.rand(1, generator=g).item()) if big else (0.15 + 0.2 * torch.rand(1, generator=g).item())
    bw, bh = side * w, side * h * (0.8 + 0.4 * torch.rand(1, generator=g).item())
    bh = min(bh, h - 2.0)
    cx = bw / 2 + torch.rand(1, generator=g).item() * (w - bw)
    cy = bh / 2 + torch.rand(1, generator=g).item() * (h - bh)
    return torch.tensor([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], dtype=torch.float32)


def build_case(name: str):
    cfg_name, seed, n_cls, n_shot, s_sizes, q_sizes = CASES[name]
    g = torch.Generator().manual_seed(1000 + seed)
    support = []
    for c in range(n_cls):
        shots = []
        for s in range(n_shot):
            h, w = s_sizes[(c * n_shot + s) % len(s_sizes)]
            shots.append({"image": synth_image(g, h, w), "box": synth_box(g, h, w, big=(s % 2 == 1))})
        support.append(shots)
    query = [synth_image(g, h, w) for (h, w) in q_sizes]
    return cfg_name, seed, support, query


def run_reference(cfg, state, support, query):
    with contextlib.redirect_stdout(io.StringIO()):
        model = reference_loader.build_reference_model(cfg)
    W.load_into_module(model, state)
    model.eval()
    np.random.seed(0)
    out = {"raw_codes": [], "norm_codes": []}
    codes = []
    with torch.no_grad():
        for c, shots in enumerate(support):
            records = []
            for s in shots:
                h, w = s["image"].shape[-2:]
                inst = up.Instances((h, w))
                inst.gt_boxes = up.Boxes(s["box"][None])
                inst.gt_classes = torch.tensor([c])
                records.append({"image": s["image"], "instances": inst, "height": h, "width": w})
            with contextlib.redirect_stdout(io.StringIO()):
                code = model([{"support_set": records}], run_type="meta_learn_test_support")
            out["raw_codes"].append({k: v.clone() for k, v in code.items()})
            codes.append({"support_set_target": torch.tensor(c), "class_name": f"class{c}",
                          "class_code": {k: v.clone() for k, v in code.items()}})
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                codes = model(None, class_code=codes, run_type="meta_learn_normalize_code")
        except TypeError as e:   # ROIEncoder.forward has no cls_norm / class_codes keywords (reference quirk)
            out["normalize_error"] = f"TypeError: {e}"
        out["norm_codes"] = [{k: v.clone() for k, v in c["class_code"].items()} for c in codes]
        packed = MetaFCOSOracle.pack_codes(codes)  # restates format_class_codes_shared (needs pycocotools to import)
        out["packed"] = packed
        batched = [{"image": q, "height": q.shape[-2], "width": q.shape[-1]} for q in query]
        res = model(batched, class_code=packed, run_type="meta_learn_test_instance")
        out["detections"] = []
        for r in res:
            inst = r["instances"]
            out["detections"].append({"boxes": inst.pred_boxes.tensor.clone(), "scores": inst.scores.clone(),
                                      "classes": inst.pred_classes.clone(), "locations": inst.locations.clone(),
                                      "levels": inst.fpn_levels.clone()})
        # head intermediates straight from the reference head
        il = model.convert_batched_inputs_to_image_list(batched)
        feats = model.backbone(il.tensor)
        feats = [feats[f] for f in cfg.MODEL.FCOS.IN_FEATURES]
        logits, regs, ctrs, ious, _, _ = model.proposal_generator.fcos_head(feats, None, False, packed)
        out["feat_abs_mean"] = [float(f.abs().mean()) for f in feats]
        out["logits"], out["reg"], out["ctr"] = logits, regs, ctrs
    return out


def compare(tag, a, b):
    a, b = a.double(), b.double()
    if a.numel() == 0 and b.numel() == 0:
        print(f"  {tag:28s} empty")
        return 0.0
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    print(f"  {tag:28s} max_abs_err {err:.3e}  max_ref {ref:.3e}")
    return err / (ref + 1e-30)


def build_base_reduce_golden():
    """Base-class all-GT path: seeded chunk codes -> weighted accumulation (restated loop) -> the REFERENCE's own
    reduce_class_code / replace_class_code (sylph/modeling/code_generator/utils.py:376-427, imported unmodified)."""
    import copy
    import logging
    from oracle import base_codes_oracle as bo
    ns = reference_loader.load()
    logging.getLogger(ns.cg_utils.__name__).setLevel(logging.ERROR)
    g = torch.Generator().manual_seed(77)
    # 6 classes; class c has total_len boxes split into chunks of <= 10; chunks are dealt round-robin to 3 "ranks";
    # class 4 misses its last chunk (acc_weight != 1 -> rebalance), class 5 has a single full chunk
    totals = {0: 23, 1: 10, 2: 37, 3: 4, 4: 26, 5: 10}
    chunks = []
    for cid, tot in totals.items():
        n_chunks = (tot + 9) // 10
        for j in range(n_chunks):
            ln = min(10, tot - 10 * j)
            if cid == 4 and j == n_chunks - 1:
                continue
            chunks.append({"cid": cid, "len": ln, "total_len": tot, "name": f"class{cid}",
                           "code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g) * 0.3,
                                    "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}})
    perm = torch.randperm(len(chunks), generator=g).tolist()
    chunks = [chunks[i] for i in perm]
    ranks = [chunks[r::3] for r in range(3)]
    per_rank = [bo.accumulate_base_codes([copy.deepcopy(c["code"]) for c in rk], [c["cid"] for c in rk],
                                         [c["len"] for c in rk], [c["total_len"] for c in rk], [c["name"] for c in rk])
                for rk in ranks]
    gathered = [c for rk in per_rank for c in rk]
    # reduce_class_code's last log line reads result["class_code"]["cls_weight_norm"], which the shipped configs never
    # produce (KeyError at utils.py:426): the reference path only runs with that log call neutralised.
    class _Quiet:
        def info(self, *a, **k):
            pass
    real_logger = ns.cg_utils.logger
    ns.cg_utils.logger = _Quiet()
    try:
        try:
            ref_reduced = ns.cg_utils.reduce_class_code(copy.deepcopy(gathered))
            quirk = None
        except KeyError as e:   # f-string argument is evaluated even with a quiet logger
            quirk = f"KeyError {e} at utils.py:426 (log line needs cls_weight_norm)"
            patched = copy.deepcopy(gathered)
            for c in patched:
                c["class_code"]["cls_weight_norm"] = torch.zeros(1) * 0.0
            ref_reduced = ns.cg_utils.reduce_class_code(patched)
            for c in ref_reduced:
                del c["class_code"]["cls_weight_norm"]
        few_shot = [{"support_set_target": torch.tensor(cid), "class_name": f"class{cid}",
                     "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}}
                    for cid in (0, 1, 2, 3, 4, 5, 6, 7)]
        with contextlib.redirect_stdout(io.StringIO()), warnings_off():
            ref_replaced = ns.cg_utils.replace_class_code(copy.deepcopy(few_shot), copy.deepcopy(ref_reduced), "cpu")
    finally:
        ns.cg_utils.logger = real_logger
    mine = bo.reduce_class_code(copy.deepcopy(gathered))
    assert [c["support_set_target"] for c in mine] == [bo._cid(c["support_set_target"]) for c in ref_reduced]
    for a, b in zip(mine, ref_reduced):
        assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
        assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
        assert "acc_weight" not in b["class_code"]
    mine_rep = bo.replace_class_code(few_shot, mine)
    for a, b in zip(mine_rep, ref_replaced):
        assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
        assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
    print(f"[base_reduce] oracle restatement == reference reduce_class_code / replace_class_code (bit exact); quirk: {quirk}")
    golden = {"chunks_per_rank": [[{k: c[k] for k in ("cid", "len", "total_len", "name", "code")} for c in rk] for rk in ranks],
              "per_rank": per_rank, "reduced": ref_reduced, "few_shot": few_shot, "replaced": ref_replaced, "quirk": quirk,
              "torch_version": torch.__version__}
    path = os.path.join(GOLDEN_DIR, "base_reduce.pt")
    torch.save(golden, path)
    print(f"[base_reduce] wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


# ---------------------------------------------------------------------------------------------------------------
# Training forward: forward_few_shot_detector_training of the REFERENCE model in train() mode.
TRAIN_CASES = {
    # name: (config, seed, class ids, shots, query images per class, cfg overrides)
    "coco_train_2way_2shot": ("coco", 11, [7, 3], 2, 2,
                              ["MODEL.META_LEARN.SHOT", 2, "MODEL.META_LEARN.QUERY_SHOT", 2,
                               "MODEL.PROPOSAL_GENERATOR.FREEZE_BBOX_BRANCH", False, "MODEL.PROPOSAL_GENERATOR.FREEZE", False]),
    # the shipped finetune setting: box branch frozen -> only loss_fcos_cls is returned (fcos_outputs.py:87-92,626-632)
    "lvis_train_3way_1shot_cls_only": ("lvis", 12, [40, 2, 17], 1, 1,
                                       ["MODEL.META_LEARN.SHOT", 1, "MODEL.META_LEARN.QUERY_SHOT", 1]),
}


def build_train_case(name: str):
    cfg_name, seed, ids, n_shot, n_query, opts = TRAIN_CASES[name]
    g = torch.Generator().manual_seed(2000 + seed)
    sizes = [(256, 320), (240, 300), (224, 288), (200, 352)]
    items = []
    k = 0
    for ci, cid in enumerate(ids):
        sup, qry = [], []
        for s in range(n_shot):
            h, w = sizes[k % len(sizes)]; k += 1
            sup.append({"image": synth_image(g, h, w), "boxes": synth_box(g, h, w, big=(s % 2 == 1))[None],
                        "classes": torch.tensor([cid])})
        for q in range(n_query):
            h, w = sizes[k % len(sizes)]; k += 1
            # a mix of wanted classes, an unrelated class (dropped by _get_gt) and nested boxes (min-area rule);
            # the LAST query image of the episode holds no wanted class at all (the "no gt" branch)
            last = ci == len(ids) - 1 and q == n_query - 1
            boxes = [synth_box(g, h, w, big=True), synth_box(g, h, w, big=False), synth_box(g, h, w, big=False),
                     synth_box(g, h, w, big=True)]
            classes = [cid, ids[(ci + 1) % len(ids)], 59, cid]
            if last:
                classes = [58, 59, 57, 56]
            qry.append({"image": synth_image(g, h, w), "boxes": torch.stack(boxes), "classes": torch.tensor(classes)})
        items.append({"support_set": sup, "query_set": qry, "support_set_target": cid})
    return cfg_name, seed, opts, items


def to_records(items):
    """golden item -> the reference's batched_inputs (data/build.py:271-282)."""
    def rec(r):
        h, w = r["image"].shape[-2:]
        inst = up.Instances((h, w))
        inst.gt_boxes = up.Boxes(r["boxes"].clone())
        inst.gt_classes = r["classes"].clone()
        return {"image": r["image"].to(torch.float32), "instances": inst, "height": h, "width": w}
    return [{"support_set": [rec(r) for r in it["support_set"]], "query_set": [rec(r) for r in it["query_set"]],
             "support_set_target": torch.tensor(it["support_set_target"])} for it in items]


def run_reference_training(cfg, state, items):
    with contextlib.redirect_stdout(io.StringIO()):
        model = reference_loader.build_reference_model(cfg)
    W.load_into_module(model, state)
    model.train()
    np.random.seed(0)
    batched = to_records(items)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), warnings_off():
        losses = model(batched)
        # ground-truth assignment straight from the reference's FCOSOutputs
        query = [r for x in batched for r in x["query_set"]]
        targets = [x["support_set_target"] for x in batched]
        gts = model._get_gt(query, support_set_targets=targets)
        il = model.convert_batched_inputs_to_image_list(query)
        feats = model.backbone(il.tensor)
        feats = [feats[f] for f in cfg.MODEL.FCOS.IN_FEATURES]
        locations = model.proposal_generator.compute_locations(feats)
        tt = model.proposal_generator.fcos_outputs._get_ground_truth(locations, gts)
    return {"losses": {k: v.clone() for k, v in losses.items()},
            "labels": torch.cat([x.reshape(-1) for x in tt["labels"]]),
            "target_inds": torch.cat([x.reshape(-1) for x in tt["target_inds"]]),
            "reg_targets": torch.cat([x.reshape(-1, 4) for x in tt["reg_targets"]]),
            "fpn_levels": torch.cat([x.reshape(-1) for x in tt["fpn_levels"]]),
            "gt_counts": [len(g) for g in gts]}


def build_training_goldens():
    for name in TRAIN_CASES:
        cfg_name, seed, opts, items = build_train_case(name)
        cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", CONFIGS[cfg_name]), ["MODEL.DEVICE", "cpu"] + opts)
        state = W.synthetic_state_dict(cfg, seed)
        ref = run_reference_training(cfg, state, items)
        orc = build_oracle(cfg, state)
        losses, ex = orc.training_forward(to_records(items))
        print(f"[{name}] reference losses {({k: float(v) for k, v in ref['losses'].items()})}; positives "
              f"{int((ref['labels'] != 100000).sum())} of {ref['labels'].numel()}; gts per query image {ref['gt_counts']}")
        assert set(losses) == set(ref["losses"]), (set(losses), set(ref["losses"]))
        worst = 0.0
        for k in ref["losses"]:
            worst = max(worst, compare(k, losses[k].reshape(1), ref["losses"][k].reshape(1)))
        assert torch.equal(ex["labels"], ref["labels"]) and torch.equal(ex["target_inds"], ref["target_inds"])
        assert torch.equal(ex["fpn_levels"], ref["fpn_levels"])
        assert torch.equal(ex["reg_targets"], ref["reg_targets"]), (ex["reg_targets"] - ref["reg_targets"]).abs().max()
        print(f"[{name}] labels / target_inds / reg_targets bit-exact; worst loss deviation {worst:.3e}")
        golden = {"case": name, "config": CONFIGS[cfg_name], "seed": seed, "opts": opts,
                  "items": [{"support_set": [{"image": r["image"].to(torch.uint8), "boxes": r["boxes"], "classes": r["classes"]}
                                             for r in it["support_set"]],
                             "query_set": [{"image": r["image"].to(torch.uint8), "boxes": r["boxes"], "classes": r["classes"]}
                                           for r in it["query_set"]],
                             "support_set_target": it["support_set_target"]} for it in items],
                  "losses": ref["losses"], "labels": ref["labels"].to(torch.int32), "target_inds": ref["target_inds"].to(torch.int32),
                  "reg_targets": ref["reg_targets"], "fpn_levels": ref["fpn_levels"].to(torch.int8), "gt_counts": ref["gt_counts"],
                  "torch_version": torch.__version__}
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(golden, path)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


# Loss-configuration variants on the inputs of "coco_train_2way_2shot": other IoU losses, no centre sampling, a wider
# sampling radius, other focal-loss parameters.  Only the outputs are stored (inputs = the base case's).
TRAIN_VARIANTS = {
    "iou_loss": ["MODEL.FCOS.LOC_LOSS_TYPE", "iou"],
    "linear_iou_loss": ["MODEL.FCOS.LOC_LOSS_TYPE", "linear_iou"],
    "no_center_sample": ["MODEL.FCOS.CENTER_SAMPLE", False],
    "radius_2p5_sizes": ["MODEL.FCOS.POS_RADIUS", 2.5, "MODEL.FCOS.SIZES_OF_INTEREST", [32, 64, 128, 256]],
    "focal_alpha_gamma": ["MODEL.FCOS.LOSS_ALPHA", 0.4, "MODEL.FCOS.LOSS_GAMMA", 1.5],
    "focal_no_alpha": ["MODEL.FCOS.LOSS_ALPHA", -1.0],
}


def build_training_variant_goldens():
    base = "coco_train_2way_2shot"
    cfg_name, seed, opts, items = build_train_case(base)
    out = {"base_case": base, "variants": {}, "torch_version": torch.__version__}
    for vname, vopts in TRAIN_VARIANTS.items():
        cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", CONFIGS[cfg_name]),
                       ["MODEL.DEVICE", "cpu"] + opts + vopts)
        state = W.synthetic_state_dict(cfg, seed)
        ref = run_reference_training(cfg, state, items)
        losses, ex = build_oracle(cfg, state).training_forward(to_records(items))
        worst = max(abs(float(losses[k]) - float(v)) / max(abs(float(v)), 1e-12) for k, v in ref["losses"].items())
        assert set(losses) == set(ref["losses"])
        assert torch.equal(ex["labels"], ref["labels"]) and torch.equal(ex["target_inds"], ref["target_inds"])
        assert torch.equal(ex["reg_targets"], ref["reg_targets"])
        print(f"[train variant {vname}] reference losses {({k: round(float(v), 5) for k, v in ref['losses'].items()})}; "
              f"positives {int((ref['labels'] != 100000).sum())}; oracle deviation {worst:.2e}, targets bit-exact")
        out["variants"][vname] = {"opts": opts + vopts, "losses": ref["losses"], "labels": ref["labels"].to(torch.int32),
                                  "target_inds": ref["target_inds"].to(torch.int32), "reg_targets": ref["reg_targets"].to(torch.float32)}
    path = os.path.join(GOLDEN_DIR, "coco_train_variants.pt")
    torch.save(out, path)
    print(f"[train variants] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


# Backward of the training step for the code generator: the REFERENCE model's own
# `sum(model(batched).values()).backward()` (what detectron2's SimpleTrainer.run_step does), the `.grad` of every
# code-generator parameter.  The 2.4 MB convolution gradients are stored as a strided sample (every GRAD_SAMPLE_STEP-th element of
# the flattened tensor) plus float64 checksums (sum, sum of squares) of the whole tensor; small tensors in full.
GRAD_SAMPLE_STEP = 7
GRAD_CASES = ["lvis_train_3way_1shot_cls_only", "coco_train_2way_2shot", "coco_train_2way_2shot_mild"]
# "coco_train_2way_2shot" saturates the classifier with its synthetic weights (loss_fcos_cls = 92.9); the "_mild" variant is the
# same episode with conv_scale = 0.5 instead of 6 (logits of order 1, the regime of a trained model)
GRAD_STATE_OVERRIDES = {"coco_train_2way_2shot_mild": {"code_generator.code_generator_head.conv_scale.scale": 0.5}}
GRAD_BASE_CASE = {"coco_train_2way_2shot_mild": "coco_train_2way_2shot"}


def run_reference_training_grads(cfg, state, items):
    with contextlib.redirect_stdout(io.StringIO()):
        model = reference_loader.build_reference_model(cfg)
    W.load_into_module(model, state)
    model.train()
    np.random.seed(0)
    batched = to_records(items)
    with contextlib.redirect_stdout(io.StringIO()), warnings_off():
        losses = model(batched)
        sum(losses.values()).backward()
    pre = ("code_generator.", "proposal_generator.fcos_head.cls_tower.")
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if n.startswith(pre) and p.grad is not None}
    grads["__all_keys_with_grad__"] = sorted(n for n, p in model.named_parameters() if p.grad is not None)
    return {k: v.detach().clone() for k, v in losses.items()}, grads


def pack_grad(t: torch.Tensor, step: int = GRAD_SAMPLE_STEP):
    flat = t.reshape(-1)
    d = flat.double()
    rec = {"shape": tuple(t.shape), "sum": float(d.sum()), "sumsq": float((d * d).sum()), "absmax": float(d.abs().max())}
    if flat.numel() > 4096:
        rec["sample_step"] = step
        rec["sample"] = flat[::step].clone()
    else:
        rec["full"] = flat.clone()
    return rec


def build_training_grad_goldens():
    out = {"cases": {}, "torch_version": torch.__version__, "sample_step": GRAD_SAMPLE_STEP}
    for name in GRAD_CASES:
        cfg_name, seed, opts, items = build_train_case(GRAD_BASE_CASE.get(name, name))
        cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", CONFIGS[cfg_name]), ["MODEL.DEVICE", "cpu"] + opts)
        state = W.synthetic_state_dict(cfg, seed)
        for k, v in GRAD_STATE_OVERRIDES.get(name, {}).items():
            state[k] = torch.full_like(state[k], v)
        ref_losses, ref_grads = run_reference_training_grads(cfg, state, items)
        all_with_grad = ref_grads.pop("__all_keys_with_grad__")
        orc = build_oracle(cfg, state)
        losses, grads, ex = orc.training_grads(to_records(items))
        assert set(grads) == set(ref_grads), (sorted(set(grads) ^ set(ref_grads)))
        worst = 0.0
        for k in sorted(ref_grads):
            worst = max(worst, compare(k.replace("code_generator.code_generator_head.", "d ").replace("proposal_generator.fcos_head.", "d "), grads[k], ref_grads[k]))
        print(f"[{name}] oracle autograd vs reference .grad: worst relative deviation {worst:.3e}")
        assert worst < 2e-5, worst
        out["cases"][name] = {"base_case": GRAD_BASE_CASE.get(name, name), "state_overrides": GRAD_STATE_OVERRIDES.get(name, {}),
                              "reference_keys_with_grad": all_with_grad,     # EVERY parameter the reference's backward reaches
                              "losses": ref_losses, "grads": {k: pack_grad(v, GRAD_SAMPLE_STEP if k.startswith("code_generator.") else 23)
                                                                for k, v in ref_grads.items()},
                              "grad_codes": {k: v.clone() for k, v in ex["grad_codes"].items()},
                              "grad_codes_source": "oracle autograd (the reference does not expose the codes); pinned through the parameter gradients"}
    path = os.path.join(GOLDEN_DIR, "train_grads.pt")
    torch.save(out, path)
    print(f"[train grads] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def base_detector_state(cfg, seed):
    """Synthetic weights of the NON-episodic model with a class-logits convolution strong enough to fire (the synthetic
    N(0, 0.01) initialiser leaves every logit at the prior)."""
    state = W.synthetic_state_dict(cfg, seed)
    k = "proposal_generator.fcos_head.cls_logits.weight"
    g = torch.Generator().manual_seed(4242 + seed)
    state[k] = torch.nn.functional.normalize(torch.randn(state[k].shape, generator=g), dim=1) * 5.0
    state["proposal_generator.fcos_head.cls_logits.bias"] = torch.randn(state[k].shape[0], generator=g) * 0.3 - 4.2
    return state


def build_base_detector_golden():
    """`run_type=None` on the reference's NON-episodic model (Meta-FCOS-pretrain.yaml: base detector inference,
    meta_one_stage_detector.py:298-323, 435-441): the reference model's own detections and head outputs."""
    name, seed = "coco_base_detector", 12
    cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", "COCO-Detection/Meta-FCOS/Meta-FCOS-pretrain.yaml"),
                   ["MODEL.DEVICE", "cpu"])
    state = base_detector_state(cfg, seed)
    g = torch.Generator().manual_seed(2000 + seed)
    query = [synth_image(g, 224, 288), synth_image(g, 200, 256)]
    with contextlib.redirect_stdout(io.StringIO()):
        model = reference_loader.build_reference_model(cfg)
    W.load_into_module(model, state)
    model.eval()
    batched = [{"image": q, "height": q.shape[-2], "width": q.shape[-1]} for q in query]
    with torch.no_grad():
        res = model(batched)                                   # run_type=None
        il = model.convert_batched_inputs_to_image_list(batched)
        feats = model.backbone(il.tensor)
        feats = [feats[f] for f in cfg.MODEL.FCOS.IN_FEATURES]
        logits, regs, ctrs, ious, _, _ = model.proposal_generator.fcos_head(feats, None, False, None)
    dets = [{"boxes": r["instances"].pred_boxes.tensor.clone(), "scores": r["instances"].scores.clone(),
             "classes": r["instances"].pred_classes.clone(), "locations": r["instances"].locations.clone(),
             "levels": r["instances"].fpn_levels.clone()} for r in res]
    print(f"[{name}] reference detections per image: {[int(d['scores'].numel()) for d in dets]}")
    orc = build_oracle(cfg, state)
    codes = {"cls_conv": state["proposal_generator.fcos_head.cls_logits.weight"], "cls_bias": state["proposal_generator.fcos_head.cls_logits.bias"]}
    mine, inter = orc.detect(query, codes, return_intermediate=True)
    worst = 0.0
    for l in range(5):
        worst = max(worst, compare(f"logits p{l + 3}", inter["logits"][l], logits[l]))
    for i, (d, r) in enumerate(zip(mine, dets)):
        assert d["scores"].numel() == r["scores"].numel() and torch.equal(d["classes"], r["classes"]) and torch.equal(d["locations"], r["locations"])
        worst = max(worst, compare(f"det boxes[{i}]", d["boxes"], r["boxes"]), compare(f"det scores[{i}]", d["scores"], r["scores"]))
    print(f"[{name}] worst relative deviation oracle vs reference: {worst:.3e}")
    path = os.path.join(GOLDEN_DIR, f"{name}.pt")
    torch.save({"case": name, "config": "COCO-Detection/Meta-FCOS/Meta-FCOS-pretrain.yaml", "seed": seed,
                "query": [q.to(torch.uint8) for q in query], "detections": dets, "logits": [t.clone() for t in logits],
                "reg": [t.clone() for t in regs], "ctr": [t.clone() for t in ctrs], "torch_version": torch.__version__}, path)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


@contextlib.contextmanager
def warnings_off():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if "--base-detector-only" in sys.argv:
        build_base_detector_golden()
        return
    if "--base-only" in sys.argv or "--all" in sys.argv or len(sys.argv) == 1:
        build_base_reduce_golden()
    if "--base-only" in sys.argv:
        return
    if "--base-detector-only" in sys.argv or "--all" in sys.argv or len(sys.argv) == 1:
        build_base_detector_golden()
    if "--base-detector-only" in sys.argv:
        return
    if "--train-variants-only" in sys.argv:
        build_training_variant_goldens()
        return
    if "--train-grads-only" in sys.argv:
        build_training_grad_goldens()
        return
    if "--train-only" in sys.argv or "--all" in sys.argv or len(sys.argv) == 1:
        build_training_goldens()
        build_training_variant_goldens()
        build_training_grad_goldens()
    if "--train-only" in sys.argv:
        return
    for name in CASES:
        cfg_name, seed, support, query = build_case(name)
        if "--only" in sys.argv and name != sys.argv[sys.argv.index("--only") + 1]:
            continue
        cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", CONFIGS[cfg_name]),
                       ["MODEL.DEVICE", "cpu"] + CASE_OPTS.get(name, []))
        state = W.synthetic_state_dict(cfg, seed)
        ref = run_reference(cfg, state, support, query)
        n_det = [int(d["scores"].numel()) for d in ref["detections"]]
        print(f"[{name}] reference detections per image: {n_det}; feature |mean| {ref['feat_abs_mean']}")
        # oracle on the same inputs
        orc = build_oracle(cfg, state)
        worst = 0.0
        raw = []
        for c, shots in enumerate(support):
            code = orc.class_code([s["image"] for s in shots], torch.stack([s["box"] for s in shots]))
            raw.append(code)
            worst = max(worst, compare(f"raw cls_conv[{c}]", code["cls_conv"], ref["raw_codes"][c]["cls_conv"]))
            worst = max(worst, compare(f"raw cls_bias[{c}]", code["cls_bias"], ref["raw_codes"][c]["cls_bias"]))
        normed = []
        for c, code in enumerate(raw):
            w, b = orc.normalize_code(code["cls_conv"], code["cls_bias"])
            normed.append({"support_set_target": c, "class_code": {"cls_conv": w, "cls_bias": b}})
            worst = max(worst, compare(f"norm cls_conv[{c}]", w, ref["norm_codes"][c]["cls_conv"]))
            worst = max(worst, compare(f"norm cls_bias[{c}]", b, ref["norm_codes"][c]["cls_bias"]))
        packed = orc.pack_codes(normed)
        dets, inter = orc.detect(query, packed, return_intermediate=True)
        for l in range(5):
            worst = max(worst, compare(f"logits p{l + 3}", inter["logits"][l], ref["logits"][l]))
            worst = max(worst, compare(f"reg p{l + 3}", inter["reg"][l], ref["reg"][l]))
        for i, d in enumerate(dets):
            r = ref["detections"][i]
            same_n = d["scores"].numel() == r["scores"].numel()
            print(f"  image {i}: oracle {d['scores'].numel()} vs reference {r['scores'].numel()} detections")
            if same_n and d["scores"].numel():
                # order: both come out of NMS in descending-score order
                worst = max(worst, compare(f"det boxes[{i}]", d["boxes"], r["boxes"]))
                worst = max(worst, compare(f"det scores[{i}]", d["scores"], r["scores"]))
                assert torch.equal(d["classes"], r["classes"]) and torch.equal(d["levels"], r["levels"])
                assert torch.equal(d["locations"], r["locations"])
        print(f"[{name}] worst relative deviation oracle vs reference: {worst:.3e}")
        golden = {
            "case": name, "config": CONFIGS[cfg_name], "seed": seed, "opts": CASE_OPTS.get(name, []),
            "normalize_error": ref.get("normalize_error"),
            "support": [[{"image": s["image"].to(torch.uint8), "box": s["box"]} for s in shots] for shots in support],
            "query": [q.to(torch.uint8) for q in query],
            "raw_codes": ref["raw_codes"], "norm_codes": ref["norm_codes"], "packed": ref["packed"],
            "detections": ref["detections"],
            "logits": [t.clone() for t in ref["logits"]], "reg": [t.clone() for t in ref["reg"]],
            "ctr": [t.clone() for t in ref["ctr"]],
            "torch_version": torch.__version__,
        }
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(golden, path)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
