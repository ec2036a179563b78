#!/usr/bin/env python
"""BASELINE.json configs[3] measured for real: a 20-way 5-shot COCO-novel episode with 8 query images, the classes and
the query images sharded contiguously over the ranks (InferenceSampler semantics) and ONE NCCL all-gather of the class
codes in between (runner.run_episode).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_sharded.py [--steps 10] [--out profiles/rNN_cfg4_nN.json]

Timing: CUDA events on every rank around `steps` episodes after `warmup`, bracketed by barrier + synchronize; the
reported time is the MAX over ranks.  Inputs are resident uint8 device tensors (synthetic, seed 1234 + 4).  Also times
the all-gather alone (the fixed-stride (classes, 257) fp32 buffer of gather_class_code) and, on one rank, the episodic
TRAINING forward (losses) of a 5-way 5-shot batch with 8 query images."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--way", type=int, default=20)
    ap.add_argument("--shot", type=int, default=5)
    ap.add_argument("--queries", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runner import gather_class_code, run_episode, shard_range
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    from tools.bench_configs import boxes as synth_boxes

    cfg = coco_meta_fcos_cfg()
    model = build_model(cfg)
    model.pixel_mean = model.pixel_mean.to(dev)
    model.load_state_dict(W.synthetic_state_dict(cfg, 0))
    h, w = 800, 1333
    g = torch.Generator().manual_seed(1234 + 4)

    def image():
        return torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)

    # every rank draws the same episode and keeps only its shards on the device
    my_cls = set(shard_range(args.way, world, rank))
    my_q = set(shard_range(args.queries, world, rank))
    bx = synth_boxes(args.way * args.shot, 1234 + 5, h, w)
    support = []
    for c in range(args.way):
        recs = []
        for s in range(args.shot):
            im = image()
            inst = Instances((h, w))
            inst.gt_boxes = Boxes(bx[c * args.shot + s][None])
            inst.gt_classes = torch.tensor([c])
            recs.append({"image": im.to(dev) if c in my_cls else im[:, :1, :1], "instances": inst, "height": h, "width": w})
        support.append({"support_set": recs, "support_set_target": torch.tensor(c), "class_name": f"class{c}"})
    query = []
    for i in range(args.queries):
        im = image()
        query.append({"image": im.to(dev) if i in my_q else im[:, :1, :1], "height": h, "width": w})

    def step():
        return run_episode(model, support, query)

    # throughput mode: the detections of episode i travel to pinned memory asynchronously and are read on the host
    # while episode i+1 is already enqueued (runner.EpisodeFuture) -- no host synchronisation between episodes
    from sylph_few_shot_detection_b200.runner import EpisodeFuture
    ring, futures = [], []

    def step_async():
        dets, counts, out_sizes = run_episode(model, support, query, return_device=True)
        if len(ring) < 4:
            ring.append((torch.empty(dets.shape, dtype=dets.dtype, pin_memory=True),
                         torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True)))
        slot = ring[step_async.k % len(ring)] if len(ring) == 4 else ring[-1]
        step_async.k += 1
        slot[0].copy_(dets, non_blocking=True)
        slot[1].copy_(counts, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        futures.append(EpisodeFuture(slot[0], slot[1], out_sizes, ev))
        return futures.pop(0).result() if len(futures) > 1 else None
    step_async.k = 0

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms = timed(step, args.warmup, args.steps)
    ms_async = timed(step_async, args.warmup, args.steps)
    ms_balanced = None
    if world > 1:
        # query images to the least-loaded ranks (runner.balanced_query_assignment) instead of contiguous shards
        from sylph_few_shot_detection_b200.runner import query_indices_of_rank
        mine_b = set(query_indices_of_rank(support, args.queries, world, rank, balance_queries=True))
        g2 = torch.Generator().manual_seed(99)
        for i in mine_b - my_q:      # this rank needs images it did not keep on the device
            query[i] = dict(query[i], image=torch.randint(0, 256, (3, h, w), generator=g2, dtype=torch.uint8).to(dev))
        ms_balanced = timed(lambda: run_episode(model, support, query, balance_queries=True), args.warmup, args.steps)
    while futures:
        futures.pop(0).result()
    res = step()
    n_det = [int(len(r["instances"])) for r in res]
    out = {"config": f"{args.way}-way {args.shot}-shot COCO-novel episode, {args.queries} query images 800x1333, classes and "
                     f"queries sharded over {world} GPU(s), one NCCL all-gather of the class codes",
           "n_gpus": world, "ms_per_episode": round(ms, 3), "episodes_per_s": round(1000.0 / ms, 2),
           "ms_per_episode_balanced_queries": round(ms_balanced, 3) if ms_balanced else None,
           "ms_per_episode_async_results": round(ms_async, 3), "episodes_per_s_async_results": round(1000.0 / ms_async, 2),
           "episode_gflop": 23251, "tflops_all_gpus": round(23251 / ms, 1),
           "classes_on_rank0": len(my_cls), "queries_on_rank0": len(my_q), "detections_rank0": n_det,
           "steps": args.steps, "warmup": args.warmup, "timing": "CUDA events, max over ranks"}
    if world > 1:
        codes = [{"support_set_target": torch.tensor(c), "class_name": f"class{c}",
                  "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, device=dev), "cls_bias": torch.randn(1, 1, 1, 1, device=dev)}}
                 for c in my_cls]
        from sylph_few_shot_detection_b200.runner import gather_class_code_known_shards
        counts = [len(shard_range(args.way, world, r)) for r in range(world)]
        meta = [(torch.tensor(c), f"class{c}") for c in range(args.way)]
        out["all_gather_general_ms"] = round(timed(lambda: gather_class_code(codes), 5, 50), 4)   # reference-shaped: 4 collectives + syncs
        out["all_gather_ms"] = round(timed(lambda: gather_class_code_known_shards(codes, counts, meta), 5, 50), 4)  # run_episode's path
        out["all_gather_bytes"] = args.way * 257 * 4
        # the same step with the normalisation fused into a peer-memory all-gather (sylph_normalize_codes_exchange):
        # exchange alone (against all-gather + normalisation of all classes), then the whole episode
        from sylph_few_shot_detection_b200.runner import exchange_codes_peer, inference_normalization
        try:
            out["nccl_gather_plus_normalize_ms"] = round(timed(lambda: inference_normalization(
                model, gather_class_code_known_shards(codes, counts, meta)), 5, 50), 4)
            out["peer_exchange_ms"] = round(timed(lambda: exchange_codes_peer(model, codes, counts, meta), 5, 50), 4)
            out["ms_per_episode_peer_exchange"] = round(timed(lambda: run_episode(model, support, query, exchange="peer"),
                                                              args.warmup, args.steps), 3)
            if ms_balanced:
                out["ms_per_episode_balanced_peer_exchange"] = round(timed(
                    lambda: run_episode(model, support, query, balance_queries=True, exchange="peer"), args.warmup, args.steps), 3)
            out["peer_exchange_timed_out"] = model.engine.exchange_status()[0]
            model.engine.exchange_teardown(None)
        except RuntimeError as e:   # e.g. no peer access between the GPUs of this box
            out["peer_exchange_error"] = str(e)[:300]
    if rank == 0:
        # ---- episodic training forward (losses only): 5-way 5-shot, 8 query images with 4 ground truths each
        tcfg = coco_meta_fcos_cfg(["MODEL.META_LEARN.SHOT", 5, "MODEL.PROPOSAL_GENERATOR.FREEZE_BBOX_BRANCH", False,
                                   "MODEL.PROPOSAL_GENERATOR.FREEZE", False])
        tm = build_model(tcfg)
        tm.pixel_mean = tm.pixel_mean.to(dev)
        tm.load_state_dict(W.synthetic_state_dict(tcfg, 0))
        tm.train()
        qb = synth_boxes(8 * 4, 77, h, w)
        items = []
        qi = 0
        for c in range(5):
            sup = []
            for s in range(5):
                inst = Instances((h, w))
                inst.gt_boxes = Boxes(bx[c * 5 + s][None])
                inst.gt_classes = torch.tensor([c])
                sup.append({"image": image().to(dev), "instances": inst, "height": h, "width": w})
            qs = []
            for _ in range(2 if c < 3 else 1):
                inst = Instances((h, w))
                inst.gt_boxes = Boxes(qb[4 * qi:4 * qi + 4])
                inst.gt_classes = torch.tensor([c, (c + 1) % 5, 50, c])
                qs.append({"image": image().to(dev), "instances": inst, "height": h, "width": w})
                qi += 1
            items.append({"support_set": sup, "query_set": qs, "support_set_target": torch.tensor(c)})
        # single-rank timing (no collectives inside: the other ranks do not take part)
        if world == 1:
            for _ in range(2):
                tm(items)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                losses = tm(items)
            e1.record()
            torch.cuda.synchronize()
            eng = tm.engine
            eng.set_profiling(True)
            tm(items)
            torch.cuda.synchronize()
            loss_ms = [t for n, t, _, _ in eng.timings() if n == "loss.targets+sums"]
            eng.set_profiling(False)
            out["training_forward"] = {"config": "5-way 5-shot, 8 query images 800x1333 with 4 ground truths each (losses only)",
                                       "ms_per_batch": round(e0.elapsed_time(e1) / 5, 3),
                                       "loss_kernel_ms": round(sum(loss_ms), 4),
                                       "losses": {k: round(float(v), 5) for k, v in losses.items()}}
        print(json.dumps(out))
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(out, indent=1) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
