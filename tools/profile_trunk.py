#!/usr/bin/env python
"""Small driver for ncu captures of the trunk kernels in the default (exact) precision mode: N passes of the backbone + FPN over
33 uint8 images of 800x1333 (the 25 support + 8 query images of the headline episode share one trunk pass).
Per pass the single-CTA staged split kernel `conv_gemm_f16_kernel<128,3,2,0,...>` runs 8 times (res2 shortcut, res2 conv3 x3,
res3 conv3 x4) and `conv1x1_pair_split_kernel` 21 times (res3.sc, res4.sc, then conv1 / conv3 of the res4 and res5 blocks)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sylph_few_shot_detection_b200 import weights as W  # noqa: E402
from sylph_few_shot_detection_b200.modeling import build_model  # noqa: E402
from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg  # noqa: E402
from sylph_few_shot_detection_b200.runtime import SLOT_QUERY  # noqa: E402

n_pass = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 33
cfg = coco_meta_fcos_cfg()
model = build_model(cfg)
model.load_state_dict(W.synthetic_state_dict(cfg, 0))
g = torch.Generator().manual_seed(1)
imgs = [torch.randint(0, 256, (3, 800, 1333), generator=g, dtype=torch.uint8).cuda() for _ in range(n_img)]
eng = model.engine
for _ in range(n_pass):
    l0 = eng.launch_count()
    eng.extract_features(SLOT_QUERY, imgs)
    torch.cuda.synchronize()
print("launches per pass:", eng.launch_count() - l0, "precision:", eng.precision if hasattr(eng, "precision") else "?")
