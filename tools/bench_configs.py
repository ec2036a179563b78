#!/usr/bin/env python
"""Timing of the other BASELINE.json configs on one GPU (the headline config lives in bench.py):

  cfg3  5-way 10-shot Meta-FCOS R-101, LVIS-shaped (POST_NMS_TOPK 300, BIAS_L2_NORM), 8 queries
  cfg4  20-way 5-shot COCO-novel episode, the per-GPU critical path at W=8 (3 classes = 15 support images, all
        20 codes, 1 query image) and the whole episode on one GPU
  cfg5  LVIS 1203-class code-generation sweep (10 shots, 12 030 ROIs) over a pool of support-image features

    python tools/bench_configs.py [--out profiles/rNN_configs.json]
"""
import argparse
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def build(cfg, seed=0):
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    model = build_model(cfg)
    model.load_state_dict(W.synthetic_state_dict(cfg, seed))
    return model


def images(n, seed, h=800, w=1333):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8).cuda() for _ in range(n)]


def boxes(n, seed, h=800, w=1333):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        side = float(torch.exp(torch.empty(1).uniform_(3.4657, 6.9078, generator=g)))
        bw, bh = min(side, w - 1.0), min(side, h - 1.0)
        cx = float(torch.empty(1).uniform_(bw / 2, w - bw / 2, generator=g))
        cy = float(torch.empty(1).uniform_(bh / 2, h - bh / 2, generator=g))
        out.append([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2])
    return torch.tensor(out)


def episode(model, sup, bx, n_way, n_shot, qry):
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    eng = model.engine
    offsets = list(range(0, n_way * n_shot + 1, n_shot))

    def run():
        eng.extract_features(SLOT_SUPPORT, sup)
        raw = eng.generate_codes(SLOT_SUPPORT, bx, list(range(len(sup))), offsets)
        codes = eng.normalize_codes(raw)
        eng.extract_features(SLOT_QUERY, qry)
        return eng.detect(SLOT_QUERY, codes, max_dets=max(2 * eng.post_nms_topk, 128))
    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg, lvis_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, SLOT_SUPPORT
    res = {}

    # ---- cfg3
    model = build(lvis_meta_fcos_cfg(["MODEL.RESNETS.DEPTH", 101]))
    ms = timed(episode(model, images(50, 1), boxes(50, 2), 5, 10, images(8, 3)), warm=2, reps=3)
    res["cfg3_5way_10shot_R101_lvis_8q"] = {"ms_per_episode": round(ms, 3), "episodes_per_s": round(1000 / ms, 2),
                                            "episode_gflop": 22497, "tflops": round(22497 / ms, 1)}
    del model
    torch.cuda.empty_cache()

    # ---- cfg4
    model = build(coco_meta_fcos_cfg())
    ms_all = timed(episode(model, images(100, 4), boxes(100, 5), 20, 5, images(8, 6)), warm=1, reps=3)
    eng = model.engine
    sup15, bx15, q1 = images(15, 7), boxes(15, 8), images(1, 9)
    codes20 = torch.randn(20, 257, device="cuda") * 0.05

    def critical_path():
        eng.extract_features(SLOT_SUPPORT, sup15)
        raw = eng.generate_codes(SLOT_SUPPORT, bx15, list(range(15)), [0, 5, 10, 15])
        codes20[:3] = eng.normalize_codes(raw)  # stands for the all-gather landing the other ranks' codes
        eng.extract_features(SLOT_QUERY, q1)
        return eng.detect(SLOT_QUERY, codes20)
    ms_cp = timed(critical_path, warm=2, reps=5)
    res["cfg4_20way_5shot_8q"] = {"one_gpu_ms_per_episode": round(ms_all, 3), "one_gpu_tflops": round(23251 / ms_all, 1),
                                  "w8_critical_path_ms_per_rank": round(ms_cp, 3),
                                  "note": "critical path = 15 support images + 1 query image on the busiest rank; the code "
                                          "all-gather (20.6 KB) is latency-bound (~20 us NCCL) and not included"}

    # ---- cfg5: features of a pool of 16 support images, 12 030 ROIs spread over them
    pool = images(16, 10)
    eng.extract_features(SLOT_SUPPORT, pool)
    n_cls, shots = 1203, 10
    bx = boxes(n_cls * shots, 11)
    roi_image = [i % 16 for i in range(n_cls * shots)]
    offsets = list(range(0, n_cls * shots + 1, shots))

    def sweep():
        raw = eng.generate_codes(SLOT_SUPPORT, bx, roi_image, offsets)
        return eng.normalize_codes(raw)
    ms5 = timed(sweep, warm=1, reps=3)
    res["cfg5_lvis_1203_class_sweep"] = {"ms": round(ms5, 3), "classes_per_s": round(1203 / (ms5 * 1e-3)),
                                         "rois": n_cls * shots, "feature_pool_images": 16,
                                         "gflop": round(0.1736 * n_cls * shots, 1),
                                         "tflops": round(0.1736 * n_cls * shots / ms5, 1)}
    del model, eng
    torch.cuda.empty_cache()

    # ---- cfg5 with the ROIEncoder generator ("roi_encoder only"): same 12 030 ROIs, EVAL_SHOT = 10
    from sylph_few_shot_detection_b200.presets import lvis_roi_encoder_cfg
    model = build(lvis_roi_encoder_cfg())
    eng = model.engine
    eng.extract_features(SLOT_SUPPORT, pool)

    def sweep_re():
        return eng.generate_codes(SLOT_SUPPORT, bx, roi_image, offsets)
    ms5r = timed(sweep_re, warm=1, reps=3)
    # per ROI: pool conv + 2 tokenizer convs 3 x 57.8 MMAC, fc1 3.2 MMAC, dense layers ~1.4 MMAC, MS_CAM 1.6 MMAC
    gflop_roi = 2e-3 * (3 * 57.8 + 3.2 + 1.4 + 1.6)
    res["cfg5_lvis_1203_class_sweep_roi_encoder"] = {"ms": round(ms5r, 3), "classes_per_s": round(1203 / (ms5r * 1e-3)),
                                                     "rois": n_cls * shots, "feature_pool_images": 16,
                                                     "gflop": round(gflop_roi * n_cls * shots, 1),
                                                     "tflops": round(gflop_roi * n_cls * shots / ms5r, 1)}
    eng.set_profiling(True)
    sweep_re()
    agg = {}
    for name, t_ms, fl, by in eng.timings():
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t_ms
    eng.set_profiling(False)
    res["cfg5_lvis_1203_class_sweep_roi_encoder"]["per_kernel_ms"] = {k: [v[0], round(v[1], 3)] for k, v in
                                                                      sorted(agg.items(), key=lambda kv: -kv[1][1])}
    out = json.dumps(res, indent=1)
    print(out)
    if args.out:
        with open(args.out, "w") as f:
            f.write(out + "\n")


if __name__ == "__main__":
    main()
