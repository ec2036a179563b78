#!/usr/bin/env python
"""Benchmark of the Meta-FCOS few-shot inference path (BASELINE.json metric: episodes/sec, 5-way 5-shot R-50).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one episode of configs[1]: 5 classes x 5 support images -> 5 class codes (backbone, ROIAlign, code
generator, normalisation), then 8 query images 800x1333 -> detections (backbone, FCOS head with the code-conditioned
classifier, proposals, NMS).  Synthetic uint8 images, synthetic "trained-like" weights (see weights.py).

  value     episodes/s with all inputs already resident in HBM, timed with CUDA events (max over ranks)
  e2e       the same through the public plugin API (MetaOneStageDetector.forward with run_type=...), inputs in pinned
            HOST memory: H2D of every image and D2H of the detections inside the timed region
  roofline  the kernel with the largest share of the step (staged 1x1 bottleneck conv, HBM-bound) and, as
            roofline_tensor, the tensor-bound FCOS tower kernel; both timed live with CUDA events
  cpu_baseline  the CPU oracle (port of the reference forward) on this box's host cores, bounded sample (rank 0, N=1)

N > 1 (torchrun, one rank per GPU): "replicas" -- every rank runs whole episodes (5 classes < 8 GPUs; the
class-sharded episode with the NCCL code all-gather is exercised by tests/test_runner_dist.py and reported under
"sharded" for the 20-way config).  Weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_WAY, N_SHOT, N_QUERY, IMG_H, IMG_W = 5, 5, 8, 800, 1333
TOWER_FLOP_PER_IMAGE_LAYER = 2.0 * 22400 * 256 * 2304      # algorithmic: 22 400 locations x Cout 256 x K 2304
EPISODE_GFLOP = 8281.0


def synth_episode(seed: int, n_way=N_WAY, n_shot=N_SHOT, n_query=N_QUERY, h=IMG_H, w=IMG_W, pinned=False):
    """uniform uint8 images, one box per support image with sqrt(area) log-uniform in [32, 1000]."""
    g = torch.Generator().manual_seed(1234 + seed)
    def img():
        t = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)
        return t.pin_memory() if pinned else t
    support, boxes = [], []
    for _ in range(n_way * n_shot):
        support.append(img())
        side = float(torch.exp(torch.empty(1).uniform_(3.4657, 6.9078, generator=g)))   # ln 32 .. ln 1000
        aspect = float(torch.exp(torch.empty(1).uniform_(-0.6931, 0.6931, generator=g)))
        bw, bh = min(side * aspect ** 0.5, w - 1.0), min(side / aspect ** 0.5, h - 1.0)
        bw, bh = max(bw, 8.0), max(bh, 8.0)
        cx = float(torch.empty(1).uniform_(bw / 2, w - bw / 2, generator=g))
        cy = float(torch.empty(1).uniform_(bh / 2, h - bh / 2, generator=g))
        boxes.append([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2])
    query = [img() for _ in range(n_query)]
    return support, torch.tensor(boxes, dtype=torch.float32), query


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_sample(cfg, state, threads: int):
    """The reference forward on the host cores (oracle port), bounded sample: ONE support image -> class code and
    ONE query image -> detections at full 800x1333 resolution; the episode time is 25 x support + 8 x query, exactly
    how the reference's batch-1 loops scale (meta_learn_evaluation.py:299,413)."""
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    torch.set_num_threads(threads)
    orc = MetaFCOSOracle(cfg, state)
    support, boxes, query = synth_episode(0, n_way=1, n_shot=1, n_query=1)
    code = orc.class_code([support[0].float()], boxes[:1])          # warm-up (thread pool, oneDNN primitives)
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        code = orc.class_code([support[0].float()], boxes[:1])
    t_support = (time.perf_counter() - t0) / reps
    w, b = orc.normalize_code(code["cls_conv"], code["cls_bias"])
    codes = {"cls_conv": w.repeat(N_WAY, 1, 1, 1), "cls_bias": b.repeat(N_WAY)}
    orc.detect([query[0].float()], codes)                          # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.detect([query[0].float()], codes)
    t_query = (time.perf_counter() - t0) / reps
    episode_s = N_WAY * N_SHOT * t_support + N_QUERY * t_query
    return {"value": 1.0 / episode_s, "unit": "episodes/s", "cores": threads, "kind": "port",
            "sample": f"1 support image ({t_support:.2f} s) + 1 query image ({t_query:.2f} s) at 800x1333, fp32, "
                      f"scaled to 25 support + 8 query images per episode (batch-1 loops like the reference)",
            "episode_seconds": episode_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    cfg = coco_meta_fcos_cfg()
    config = {"workload": "5-way 5-shot Meta-FCOS R-50 FPN, 8 query images 800x1333 (configs[1])", "n_way": N_WAY,
              "n_shot": N_SHOT, "n_query": N_QUERY, "image": [IMG_H, IMG_W], "parallelism": f"replicas x{world}",
              "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        state = W.synthetic_state_dict(cfg, 0)
        threads = os.cpu_count() or 1
        vals = []
        for i in range(args.warmup + args.steps):
            r = cpu_reference_sample(cfg, state, threads)
            if i >= args.warmup:
                vals.append(r)
        v = sum(x["value"] for x in vals) / max(len(vals), 1)
        last = vals[-1]
        last["value"] = v
        print(json.dumps({"impl": "reference", "metric": "episodes/sec 5-way 5-shot Meta-FCOS R-50", "value": v,
                          "unit": "episodes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": last,
                          "e2e": {"value": v, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    state = W.synthetic_state_dict(cfg, 0)
    model = build_model(cfg)
    model.pixel_mean = model.pixel_mean.to(dev)
    model.load_state_dict(state)
    eng = model.engine

    support_h, boxes, query_h = synth_episode(rank, pinned=True)
    support_d = [t.to(dev) for t in support_h]
    query_d = [t.to(dev) for t in query_h]
    offsets = list(range(0, N_WAY * N_SHOT + 1, N_SHOT))
    roi_image = list(range(N_WAY * N_SHOT))

    merged_trunk = os.environ.get("SYLPH_BENCH_SPLIT_TRUNK", "0") != "1"

    def episode_device():
        if merged_trunk:   # support + query batches through one bottom-up trunk pass, FPN per slot
            eng.extract_features_multi([(SLOT_SUPPORT, support_d), (SLOT_QUERY, query_d)])
        else:
            eng.extract_features(SLOT_SUPPORT, support_d)
        if merged_trunk:   # codes on the engine's side stream while the towers run (SYLPH_OVERLAP_CODEGEN=0: one stream)
            return eng.generate_and_detect(SLOT_SUPPORT, SLOT_QUERY, boxes, roi_image, offsets)[0]
        raw = eng.generate_codes(SLOT_SUPPORT, boxes, roi_image, offsets)
        codes = eng.normalize_codes(raw)
        eng.extract_features(SLOT_QUERY, query_d)
        return eng.detect(SLOT_QUERY, codes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, out

    for _ in range(max(args.warmup, 3)):
        episode_device()
    # the same episode captured once into a CUDA graph (runner.EpisodeGraph): one graph launch per step instead of
    # ~125 kernel launches; inputs are refreshed in the static buffers inside the timed region (device-to-device)
    use_graph = os.environ.get("SYLPH_BENCH_GRAPH", "1") != "0"
    graph = None
    if use_graph:
        from sylph_few_shot_detection_b200.runner import EpisodeGraph
        graph = EpisodeGraph(model, N_WAY, N_SHOT, N_QUERY, (IMG_H, IMG_W))
        boxes_d = boxes.to(dev)
        l_before = eng.launch_count()
        episode_device()
        launches_per_episode = eng.launch_count() - l_before

        def episode_graph():
            for dst, src in zip(graph.support, support_d):
                dst.copy_(src, non_blocking=True)
            for dst, src in zip(graph.query, query_d):
                dst.copy_(src, non_blocking=True)
            graph.boxes.copy_(boxes_d, non_blocking=True)
            return graph.replay()
        for _ in range(3):
            episode_graph()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launch_count()
    ms_eager, (dets, counts) = timed(episode_device, args.steps)
    launches = eng.launch_count() - l0
    ms = ms_eager
    if use_graph:
        ms_graph, (dets_g, counts_g) = timed(episode_graph, args.steps)
        config["launch"] = {"eager_ms_per_step": round(ms_eager / args.steps, 3), "cuda_graph_ms_per_step": round(ms_graph / args.steps, 3),
                            "value_from": "cuda_graph" if ms_graph < ms_eager else "eager"}
        if ms_graph < ms_eager:
            ms, dets, counts = ms_graph, dets_g, counts_g
            launches = launches_per_episode * args.steps   # kernels inside the replayed graphs
    clocks = sampler.stop() if sampler else None
    value = world * args.steps / (ms / 1000.0)

    # ---- end to end through the plugin API, host-resident inputs
    support_items = []
    for c in range(N_WAY):
        recs = []
        for s in range(N_SHOT):
            i = c * N_SHOT + s
            inst = Instances((IMG_H, IMG_W))
            inst.gt_boxes = Boxes(boxes[i:i + 1])
            inst.gt_classes = torch.tensor([c])
            recs.append({"image": support_h[i], "instances": inst, "height": IMG_H, "width": IMG_W})
        support_items.append({"support_set": recs, "support_set_target": torch.tensor(c), "class_name": f"class{c}"})
    query_items = [{"image": q, "height": IMG_H, "width": IMG_W} for q in query_h]

    # Every step copies one episode's images from pinned host memory (the NEXT episode's, double-buffered on a side
    # stream while the current one computes) and reads the current episode's detections back to the host.
    from sylph_few_shot_detection_b200.runner import EpisodePipeline
    pipe = EpisodePipeline(model)
    pending = [pipe.submit(support_items, query_items)]

    # One step = submit the H2D copies of the NEXT episode, enqueue the current one (kernels + asynchronous D2H of its
    # detections into pinned memory), then read the PREVIOUS episode's detections on the host while the device works.
    # Every step therefore moves one episode's inputs H2D and one episode's results D2H and materialises them as host
    # Instances; nothing is created on the device.  SYLPH_BENCH_E2E_SYNC=1 reads each episode's results before the
    # next one is enqueued (the device then idles while the host unpacks / submits: 0.7 ms per step).
    in_flight = []

    def episode_e2e():
        pending.append(pipe.submit(support_items, query_items))
        if os.environ.get("SYLPH_BENCH_E2E_SYNC", "0") == "1":
            res = pipe.run(pending.pop(0))
            return [(r["instances"].pred_boxes.tensor.cpu(), r["instances"].scores.cpu()) for r in res]
        in_flight.append(pipe.run_async(pending.pop(0)))
        # the previous episode's detections (the very first call has no predecessor and reads its own)
        res = in_flight.pop(0).result() if len(in_flight) > 1 else in_flight[0].result()
        return [(r["instances"].pred_boxes.tensor, r["instances"].scores) for r in res]

    for _ in range(3):
        episode_e2e()
    # diagnostic: pinned-host -> device bandwidth of one episode's images on this box / NUMA placement (when it drops
    # below h2d_bytes_per_step / ms_per_step the end-to-end number is bound by the copy, not by the kernels)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    _tmp = [t.to(dev, non_blocking=True) for t in support_h + query_h]
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = sum(t.numel() for t in support_h + query_h) / (c0.elapsed_time(c1) * 1e-3) * 1e-9
    pinned_ok = all(t.is_pinned() for t in support_h + query_h)
    del _tmp
    ms_e2e, res = timed(episode_e2e, args.steps)
    while in_flight:
        in_flight.pop(0).result()
    e2e_value = world * args.steps / (ms_e2e / 1000.0)
    h2d = sum(t.numel() for t in support_h + query_h) + boxes.numel() * 4
    if pipe._ring:   # asynchronous path: the whole fixed-size (images, max_dets, 9) fp32 buffer + the counts travel
        d2h = pipe._ring[0][0].numel() * 4 + pipe._ring[0][1].numel() * 4
    else:
        d2h = sum(b.numel() * 4 + s.numel() * 4 for b, s in res) + N_WAY * 257 * 4

    # ---- roofline of the dominant kernel (FCOS tower layer), live CUDA events around each launch
    roofline, roofline_tensor, breakdown = None, None, None
    if rank == 0:
        peaks = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                peaks.update(json.load(f))
                peaks["source"] = "measured"
        except Exception:
            pass
        eng.set_profiling(True)
        episode_device()
        tm = eng.timings()
        eng.set_profiling(False)
        agg = {}
        for name, t_ms, fl, by in tm:
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += t_ms; a[2] += fl; a[3] += by
        total_ms = sum(a[1] for a in agg.values())
        breakdown = {k: {"launches": a[0], "ms": round(a[1], 4), "share": round(a[1] / total_ms, 4),
                         "tflops_padded": round(a[2] / a[1] * 1e-9, 1) if a[1] > 0 else None,
                         "gbs_algorithmic": round(a[3] / a[1] * 1e-6, 1) if a[1] > 0 and a[3] > 0 else None} for k, a in
                     sorted(agg.items(), key=lambda kv: -kv[1][1])}
        traffic = {}
        try:
            with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass

        def ncu_traffic(key):
            t = traffic.get(key)
            return float(t["bytes_per_launch"]) if t else None

        # (1) the kernel with the largest share of the step: the staged-epilogue 1x1 convolution of the bottleneck
        # blocks (conv3 + residual).  HBM-bound; algorithmic bytes = interior pixels x (Cin read + residual read +
        # output write) x 2 B per launch, summed over the 16 launches of one backbone pass.
        conv3 = [t for n, t, _, _ in tm if n == "res.conv3_1x1"]
        hp, wp = (IMG_H + 31) // 32 * 32, (IMG_W + 31) // 32 * 32
        conv3_bytes_img = sum(nb * (hp >> (s + 2)) * (wp >> (s + 2)) * ((64 << s) + 2 * (256 << s)) * 2
                              for s, nb in enumerate([3, 4, 6, 3]))
        if conv3:
            tot_ms = sum(conv3)
            n_img = N_WAY * N_SHOT + N_QUERY
            achieved = conv3_bytes_img * n_img / (tot_ms * 1e-3) * 1e-9
            peak = float(peaks["hbm_gbs"])
            roofline = {"kernel": "conv_gemm_f16_kernel<256,2,2,0> -- staged TMA-in/TMA-out 1x1 conv + residual + ReLU "
                                  "(bottleneck conv3, res2..res5; the 3 res5 launches run its CTA-pair variant conv1x1_pair_staged_kernel), "
                                  "the largest share of the step",
                        "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": ncu_traffic("conv_gemm_f16_kernel<256,2,2,0>.res2_conv3"),
                        "avg_launch_ms": round(tot_ms / len(conv3), 4), "launches_timed": len(conv3),
                        "share_of_step": breakdown["res.conv3_1x1"]["share"],
                        "bytes_per_launch": conv3_bytes_img * n_img / len(conv3),
                        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})",
                        "note": "achieved = sum of algorithmic bytes / sum of launch durations over the launches of one "
                                "episode (shapes differ per stage); traffic = ncu DRAM bytes of ONE res2 launch at 8 images "
                                "(algorithmic for that launch: 8 x 67200 px x 1152 B = 619 MB)"}
        # (2) the tensor-bound kernel: CTA-pair 3x3 convolution of the FCOS towers
        tower = [t for n, t, _, _ in tm if n in ("head.cls_tower3x3", "head.bbox_tower3x3")]
        if tower:
            avg_ms = sum(tower) / len(tower)
            achieved = TOWER_FLOP_PER_IMAGE_LAYER * N_QUERY / (avg_ms * 1e-3) * 1e-12
            peak = float(peaks["bf16_tflops_sustained"])
            roofline_tensor = {"kernel": "conv3x3_pair_kernel<3,8> -- cta_group::2 halo conv (FCOS tower 3x3 256->256, all levels x 8 images)",
                               "bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                               "frac": round(achieved / peak, 4), "traffic": ncu_traffic("conv3x3_pair_kernel.tower"),
                               "avg_launch_ms": round(avg_ms, 4), "launches_timed": len(tower),
                               "share_of_step": round(breakdown["head.cls_tower3x3"]["share"] + breakdown["head.bbox_tower3x3"]["share"], 4),
                               "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); the kernel issues "
                                              f"tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate), same dense rate as bf16",
                               "flop_per_launch": TOWER_FLOP_PER_IMAGE_LAYER * N_QUERY}
            if roofline is None:
                roofline = roofline_tensor
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"per_kernel": breakdown, "episode_ms_sum_of_timed": total_ms}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(cfg, state, os.cpu_count() or 1)

    if rank == 0:
        out = {"metric": "episodes/sec 5-way 5-shot Meta-FCOS R-50", "value": round(value, 3), "unit": "episodes/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic", "config": config,
               "e2e": {"value": round(e2e_value, 3), "unit": "episodes/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "h2d_copy_gbs": round(h2d_gbs, 1), "inputs_pinned": pinned_ok, "ms_per_step": round(ms_e2e / args.steps, 3)},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_tensor": roofline_tensor, "cpu_baseline": cpu,
               "episode_tflops": round(EPISODE_GFLOP * 1e-3 * value / world, 1),
               "detections_per_image": [int(c) for c in counts.cpu().tolist()],
               "e2e_detections_per_image": [int(s.numel()) for _, s in res], "per_kernel": breakdown}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
