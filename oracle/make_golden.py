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
from sylph_few_shot_detection_b200 import weights as W  # noqa: E402
from sylph_few_shot_detection_b200.config import load_cfg  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
CONFIGS = {
    "coco": "COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml",
    "lvis": "LVISv1-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml",
}
# vendored copies of the two YAML trees are NOT kept; tests rebuild the cfg from these overrides on top of defaults
CASES = {
    # name: (config, seed, classes, shots, support (H, W) list, query (H, W) list)
    "coco_2way_2shot": ("coco", 3, 2, 2, [(256, 320), (240, 300)], [(256, 320), (200, 288)]),
    "lvis_1way_3shot": ("lvis", 5, 1, 3, [(224, 256), (256, 224), (200, 240)], [(224, 288)]),
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
        with contextlib.redirect_stdout(io.StringIO()):
            codes = model(None, class_code=codes, run_type="meta_learn_normalize_code")
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


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in CASES:
        cfg_name, seed, support, query = build_case(name)
        cfg = load_cfg(os.path.join(reference_loader.REFERENCE_ROOT, "configs", CONFIGS[cfg_name]),
                       ["MODEL.DEVICE", "cpu"])
        state = W.synthetic_state_dict(cfg, seed)
        ref = run_reference(cfg, state, support, query)
        n_det = [int(d["scores"].numel()) for d in ref["detections"]]
        print(f"[{name}] reference detections per image: {n_det}; feature |mean| {ref['feat_abs_mean']}")
        # oracle on the same inputs
        orc = MetaFCOSOracle(cfg, state)
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
            "case": name, "config": CONFIGS[cfg_name], "seed": seed,
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
