#!/usr/bin/env python
"""BASELINE.json configs[4] on N GPUs: the LVIS 1203-class code-generation sweep (10 shots, 12 030 ROIs over a pool of 16
support-image pyramids, SURVEY.md 8d "Config 5"), class-parallel -- rank r generates the raw codes of a contiguous class
shard (InferenceSampler semantics, runner.shard_range) -- followed by the exchange that leaves the (1203, 257) NORMALISED
codes on every rank.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/bench_sweep_sharded.py [--steps 5] [--generator CodeGenerator|ROIEncoder] [--out profiles/rNN_cfg5_nN.json]

Both forms of the exchange are timed: "nccl" = one all_gather_into_tensor of the padded raw shards + normalisation of all
1203 classes on every rank (what the reference's _gather_class_code + inference_normalization amount to), and "peer" =
sylph_normalize_codes_exchange (each rank normalises its shard, the kernel's stores land in every rank's buffer over
NVLink).  CUDA events, barrier + synchronize on both sides, MAX over ranks.  Every rank checks that the two forms give
bit-identical codes and rank 0 that they equal the single-GPU sweep."""
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--classes", type=int, default=1203)
    ap.add_argument("--shots", type=int, default=10)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--generator", default="CodeGenerator", choices=["CodeGenerator", "ROIEncoder"])
    ap.add_argument("--out", default="")
    ap.add_argument("--profile", action="store_true", help="per-stage breakdown of one sweep (CUDA events around every launch)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg, lvis_roi_encoder_cfg
    from sylph_few_shot_detection_b200.runner import shard_range
    from sylph_few_shot_detection_b200.runtime import CODE_STRIDE, SLOT_SUPPORT
    from tools.bench_configs import boxes, build, images

    roi_encoder = args.generator == "ROIEncoder"
    model = build(lvis_roi_encoder_cfg() if roi_encoder else coco_meta_fcos_cfg())
    model.pixel_mean = model.pixel_mean.to(dev)
    eng = model.engine
    eng.extract_features(SLOT_SUPPORT, images(args.pool, 10))          # every rank holds the same feature pool
    n_cls, shots = args.classes, args.shots
    bx = boxes(n_cls * shots, 11)
    mine = shard_range(n_cls, world, rank)
    counts = [len(shard_range(n_cls, world, r)) for r in range(world)]
    max_n = max(counts)
    my_boxes = bx[mine.start * shots:mine.stop * shots]
    my_roi_image = [i % args.pool for i in range(mine.start * shots, mine.stop * shots)]
    my_offsets = list(range(0, len(mine) * shots + 1, shots))
    norm = (lambda t: t) if roi_encoder else eng.normalize_codes        # ROIEncoder codes are final

    def generate():
        return eng.generate_codes(SLOT_SUPPORT, my_boxes, my_roi_image, my_offsets)

    def sweep_nccl():
        raw = generate()
        if world == 1:
            return norm(raw)
        buf = raw if raw.shape[0] == max_n else torch.cat([raw, raw.new_zeros((max_n - raw.shape[0], CODE_STRIDE))])
        gathered = torch.empty((world * max_n, CODE_STRIDE), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, buf.contiguous())
        rows = gathered if all(c == max_n for c in counts) else torch.cat(
            [gathered[r * max_n:r * max_n + counts[r]] for r in range(world)])
        return norm(rows)

    def sweep_peer():
        return eng.normalize_codes_exchange(generate(), mine.start, n_cls)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    if args.profile and rank == 0:
        generate()
        eng.set_profiling(True)
        generate()
        torch.cuda.synchronize()
        agg = {}
        for name, t_ms, fl, by in eng.timings():
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1; a[1] += t_ms; a[2] += fl
        eng.set_profiling(False)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:28s} launches {a[0]:3d}  {a[1]:8.3f} ms  {a[2] / max(a[1], 1e-9) * 1e-9:8.1f} TFLOP/s executed", file=sys.stderr)
        print(f"  sum of timed stages {sum(a[1] for a in agg.values()):.3f} ms", file=sys.stderr)
    ms_gen = timed(generate)
    ms_nccl = timed(sweep_nccl)
    out = {"config": f"LVIS {n_cls}-class code-generation sweep ({args.generator}), {shots} shots = {n_cls * shots} ROIs over a pool "
                     f"of {args.pool} support pyramids, classes sharded over {world} GPU(s)",
           "n_gpus": world, "classes_on_rank0": counts[0], "ms_generate_only": round(ms_gen, 3),
           "ms_sweep_nccl_gather_then_normalize": round(ms_nccl, 3), "classes_per_s_nccl": round(n_cls / (ms_nccl * 1e-3)),
           "exchange_bytes": n_cls * CODE_STRIDE * 4, "steps": args.steps, "warmup": args.warmup,
           "timing": "CUDA events, max over ranks"}
    ref = sweep_nccl()
    try:
        eng.exchange_setup(None, max_classes=max(2048, n_cls))
        ms_peer = timed(sweep_peer)
        got = sweep_peer()
        out["ms_sweep_peer_exchange"] = round(ms_peer, 3)
        out["classes_per_s_peer"] = round(n_cls / (ms_peer * 1e-3))
        out["peer_equals_nccl_bitwise"] = bool(torch.equal(got, ref))
        out["peer_exchange_timed_out"] = eng.exchange_status()[0]
        eng.exchange_teardown(None)
    except RuntimeError as e:       # e.g. no peer access between the GPUs of this box
        out["peer_exchange_error"] = str(e)[:300]
    if world > 1:
        # the sharded sweep against the single-GPU sweep of all classes (rank 0 recomputes it)
        if rank == 0:
            raw_all = eng.generate_codes(SLOT_SUPPORT, bx, [i % args.pool for i in range(n_cls * shots)],
                                         list(range(0, n_cls * shots + 1, shots)))
            out["sharded_equals_single_gpu_bitwise"] = bool(torch.equal(norm(raw_all), ref))
        ok = torch.tensor([1 if out.get("peer_equals_nccl_bitwise", True) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        out["peer_equals_nccl_on_all_ranks"] = bool(ok.item())
    if rank == 0:
        print(json.dumps(out))
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(out, indent=1) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
