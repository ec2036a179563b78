#!/usr/bin/env python
"""Meta-training loop of the hyper-network stage on one B200, written exactly as one would write it against the reference
(`configs/COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml`: backbone and box branch frozen, code generator and FCOS class tower
trained): build the model, `train()`, an optimiser over `model.parameters()`, `sum(losses.values()).backward()`, `step()`.

    python examples/train_code_generator.py [--iters 5] [--lr 1e-4]

Synthetic weights and images (no datasets offline); prints the loss of the SAME episode batch after every step -- it goes down."""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--lr", type=float, default=1e-4)
    args = ap.parse_args()
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.structures import Boxes, Instances

    cfg = coco_meta_fcos_cfg(["MODEL.META_LEARN.SHOT", 2, "MODEL.META_LEARN.QUERY_SHOT", 1])
    model = build_model(cfg)
    model.load_state_dict(W.synthetic_state_dict(cfg, 0))
    model.train()                                            # registers code_generator.* and fcos_head.cls_tower.* as parameters
    print(f"{sum(p.numel() for p in model.parameters()) / 1e6:.2f} M trainable values in {len(list(model.parameters()))} tensors")
    g = torch.Generator().manual_seed(0)
    h, w = 320, 416

    def record(classes, boxes):
        inst = Instances((h, w))
        inst.gt_boxes = Boxes(torch.tensor(boxes, dtype=torch.float32))
        inst.gt_classes = torch.tensor(classes)
        return {"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).cuda(), "instances": inst, "height": h, "width": w}
    batched = [{"support_set": [record([c], [[40.0 + 30 * c, 50.0, 200.0 + 40 * c, 220.0]]) for _ in range(2)],
                "query_set": [record([c, (c + 1) % 3], [[60.0, 40.0 + 20 * c, 260.0, 250.0], [150.0, 100.0, 330.0, 300.0]])],
                "support_set_target": torch.tensor(c)} for c in range(3)]
    opt = torch.optim.SGD(model.parameters(), lr=args.lr)
    for it in range(args.iters):
        opt.zero_grad(set_to_none=True)
        losses = model(batched)
        total = sum(losses.values())
        total.backward()
        opt.step()
        print(f"iter {it}: " + ", ".join(f"{k} {float(v.detach()):.5f}" for k, v in losses.items()))


if __name__ == "__main__":
    main()
