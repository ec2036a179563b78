#!/usr/bin/env python
"""One meta-training iteration of the hyper-network stage on one GPU: the per-GPU batch of the shipped
meta-training configs (MODEL.META_LEARN.CLASS 3 x SHOT 5 support images + QUERY_SHOT 1 query image per class, 800x1333),
backbone and box branch frozen, code generator and FCOS class tower trained (Meta-FCOS-finetune.yaml).

    python tools/bench_training.py [--steps 10] [--classes 3] [--shot 5] [--out profiles/rNN_training_step.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_training.py   # data parallel:
        every rank trains on its own episode batch, the model is wrapped in DistributedDataParallel (gradient all-reduce over NCCL), the loss
        normalisers are summed over the ranks; time = max over ranks

Reports ms per iteration for (a) the training forward alone (losses), (b) forward + backward of the code generator
(sylph_fcos_cls_loss_backward + sylph_codegen_backward + sylph_cls_tower_backward behind `sum(losses.values()).backward()`), (c) forward + backward +
an SGD step + the weight refresh of the engine (sylph_update_code_generator), and the per-stage device times of the
backward kernels.  CUDA events, inputs resident on the device (uint8), synthetic weights and images."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--shot", type=int, default=5)
    ap.add_argument("--query-shot", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"],
                    help="operand scheme of the tensor-core kernels, forward and backward (DESIGN.md section 5)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    from tools.bench_configs import boxes as synth_boxes

    cfg = coco_meta_fcos_cfg(["MODEL.META_LEARN.SHOT", args.shot, "MODEL.META_LEARN.QUERY_SHOT", args.query_shot])
    model = build_model(cfg, args.precision)
    model.to(dev)
    model.load_state_dict(W.synthetic_state_dict(cfg, 0))
    model.train()
    h, w = 800, 1333
    g = torch.Generator().manual_seed(4321 + rank)                 # every rank its own episode batch
    bx = synth_boxes(args.classes * (args.shot + 2 * args.query_shot), 4322 + rank, h, w)
    k = 0
    batched = []
    for c in range(args.classes):
        sup, qry = [], []
        for _ in range(args.shot):
            inst = Instances((h, w))
            inst.gt_boxes = Boxes(bx[k][None]); k += 1
            inst.gt_classes = torch.tensor([c])
            sup.append({"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).to(dev), "instances": inst,
                        "height": h, "width": w})
        for _ in range(args.query_shot):
            inst = Instances((h, w))
            inst.gt_boxes = Boxes(bx[k:k + 2]); k += 2
            inst.gt_classes = torch.tensor([c, (c + 1) % args.classes])
            qry.append({"image": torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).to(dev), "instances": inst,
                        "height": h, "width": w})
        batched.append({"support_set": sup, "query_set": qry, "support_set_target": torch.tensor(c)})
    opt = torch.optim.SGD(model.parameters(), lr=1e-6)
    net = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        net = DDP(model, device_ids=[local], find_unused_parameters=True)   # DDP_FIND_UNUSED_PARAMETERS: True in the shipped configs

    def fwd():
        with torch.no_grad():
            return model(batched)

    def fwd_bwd():
        model.zero_grad(set_to_none=True)
        losses = net(batched)
        sum(losses.values()).backward()
        return losses

    def full():
        losses = fwd_bwd()
        opt.step()
        return losses

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_fwd, ms_fb, ms_full = timed(fwd), timed(fwd_bwd), timed(full)
    eng = model.engine
    eng.set_profiling(True)
    losses = fwd_bwd()
    torch.cuda.synchronize()
    stages = {}
    for name, ms, _, _ in eng.timings():
        if name.startswith("bwd.") or name.startswith("loss."):
            stages[name] = round(stages.get(name, 0.0) + ms, 4)
    eng.set_profiling(False)
    out = {"workload": f"{args.classes} classes x {args.shot} support images + {args.classes * args.query_shot} query images, 800x1333, "
                       "R-50 FPN, backbone and box branch frozen, code generator + FCOS class tower trained (per-GPU batch of the shipped meta-training configs)",
           "precision": eng.precision, "ms_forward_losses": round(ms_fwd, 3), "ms_forward_backward": round(ms_fb, 3),
           "ms_forward_backward_sgd_refresh": round(ms_full, 3), "backward_stage_ms": stages,
           "loss_fcos_cls": float(losses["loss_fcos_cls"].detach()), "steps": args.steps, "warmup": args.warmup, "n_gpus": world,
           "iterations_per_s_all_gpus": round(world * 1000.0 / ms_full, 2),
           "episode_batches": "one per rank (data parallel, DistributedDataParallel gradient all-reduce)" if world > 1 else "one"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps(out))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
