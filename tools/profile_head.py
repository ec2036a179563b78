#!/usr/bin/env python
"""Small driver for ncu captures of the FCOS-head kernels: one backbone pass over 8 query images (800x1333), then
`detect` N times.  With `-k regex:conv_gemm_f16_kernel` the backbone contributes 57 GEMM launches (R-50: stem, 52
bottleneck convs... see launch list) and every detect 10 (4 cls-tower, cond-cls, 4 bbox-tower, predictor)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sylph_few_shot_detection_b200 import weights as W  # noqa: E402
from sylph_few_shot_detection_b200.modeling import build_model  # noqa: E402
from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg  # noqa: E402
from sylph_few_shot_detection_b200.runtime import SLOT_QUERY  # noqa: E402

n_detect = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = coco_meta_fcos_cfg()
model = build_model(cfg)
model.load_state_dict(W.synthetic_state_dict(cfg, 0))
g = torch.Generator().manual_seed(1)
imgs = [torch.randint(0, 256, (3, 800, 1333), generator=g, dtype=torch.uint8).cuda() for _ in range(8)]
codes = torch.randn(5, 257, generator=g).cuda() * 0.05
codes[:, 256] = -4.0
eng = model.engine
eng.extract_features(SLOT_QUERY, imgs)
l0 = eng.launch_count()
for _ in range(n_detect):
    dets, counts = eng.detect(SLOT_QUERY, codes)
torch.cuda.synchronize()
print("launches per detect:", (eng.launch_count() - l0) // n_detect, "detections:", counts.tolist())
