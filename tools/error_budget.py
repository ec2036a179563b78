#!/usr/bin/env python
"""fp64-oracle error budget of the operand formats.

Runs the CPU oracle in float64 on one full-size image and re-runs it with the roundings the CUDA engine applies
emulated in place: every convolution's input activations and (FrozenBN-folded) weights are rounded to the operand
format of the region they belong to (trunk = ResNet bottom-up, fpn = laterals / outputs / p6 / p7, head = FCOS towers
+ predictors + the code-conditioned classifier), every stored activation (ReLU outputs, pyramid levels) is rounded to
the storage format.  Accumulation stays float64, so the numbers isolate OPERAND / STORAGE rounding from everything else.

Formats:  f16     one fp16 value (10-bit mantissa, round to nearest)           -- the "fast" mode of the engine
          f16x2   hi + lo fp16 pair (hi = rn(x), lo = rn(x - hi))               -- the "exact" mode (3 MMAs per product)
          bf16 / bf16x2, tf32t (19-bit truncation) for comparison
          f64     no rounding

Usage:  python tools/error_budget.py [--h 800 --w 1333] [--smooth] [--out profiles/r02_error_budget.md]
This is test infrastructure (it imports oracle/); nothing in the product imports it.
"""
from __future__ import annotations

import argparse
import contextlib
import os
import sys
import time

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def q_f16(x):
    return x.to(torch.float16).to(x.dtype)


def q_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def q_split(q):
    def f(x):
        hi = q(x)
        return hi + q(x - hi)
    return f


def q_tf32t(x):
    y = x.to(torch.float32).contiguous()
    bits = y.view(torch.int32) & ~0x1FFF
    return bits.view(torch.float32).to(x.dtype)


QUANT = {"f64": lambda x: x, "f16": q_f16, "f16x2": q_split(q_f16), "bf16": q_bf16, "bf16x2": q_split(q_bf16),
         "tf32t": q_tf32t}


class Emulation:
    """Patches F.conv2d / F.relu / F.relu_ so that operands and stored activations are rounded per region."""

    def __init__(self, policy):
        self.policy = policy       # region -> (activation format, weight format)
        self.region = "trunk"

    @contextlib.contextmanager
    def active(self):
        conv2d, relu, relu_ = F.conv2d, F.relu, F.relu_

        def conv(x, w, b=None, *a, **k):
            qa, qw = self.policy[self.region]
            return conv2d(QUANT[qa](x), QUANT[qw](w), b, *a, **k)

        def r(x, inplace=False):
            return QUANT[self.policy[self.region][0]](relu(x))

        def r_(x):
            return QUANT[self.policy[self.region][0]](relu(x))

        F.conv2d, F.relu, F.relu_ = conv, r, r_
        torch.relu_  # noqa: B018  (F.relu_ is what the oracle calls)
        try:
            yield self
        finally:
            F.conv2d, F.relu, F.relu_ = conv2d, relu, relu_


def run(orc, image, codes, policy):
    em = Emulation(policy)
    bu = orc.backbone.bottom_up
    h1 = bu.register_forward_pre_hook(lambda m, i: setattr(em, "region", "trunk"))
    h2 = bu.register_forward_hook(lambda m, i, o: setattr(em, "region", "fpn"))
    try:
        with em.active(), torch.no_grad():
            il = orc.preprocess([image])
            feats = orc.features(il.tensor)
            feats = [QUANT[policy["fpn"][0]](f) for f in feats]      # the pyramid is stored in the storage format
            em.region = "head"
            logits, regs, ctrs, _ = orc.head(feats, codes)
    finally:
        h1.remove()
        h2.remove()
    return feats, logits, regs, ctrs


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=800)
    ap.add_argument("--w", type=int, default=1333)
    ap.add_argument("--smooth", action="store_true", help="smooth test-style image instead of uniform noise")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    cfg = coco_meta_fcos_cfg(["MODEL.RESNETS.DEPTH", args.depth]) if args.depth != 50 else coco_meta_fcos_cfg()
    state = W.synthetic_state_dict(cfg, args.seed)
    g = torch.Generator().manual_seed(1234 + args.seed)
    if args.smooth:
        base = torch.rand(3, args.h // 8 + 2, args.w // 8 + 2, generator=g) * 255.0
        img = F.interpolate(base[None], size=(args.h, args.w), mode="bilinear", align_corners=False)[0]
        img = (img + (torch.rand(3, args.h, args.w, generator=g) - 0.5) * 40.0).clamp(0, 255).round()
    else:
        img = torch.randint(0, 256, (3, args.h, args.w), generator=g).float()
    cw = F.normalize(torch.randn(5, 256, 1, 1, generator=g), dim=1) * 5.0
    codes = {"cls_conv": cw, "cls_bias": torch.randn(5, generator=g) * 0.3 - 4.0}

    orc64 = MetaFCOSOracle(cfg, state, dtype=torch.float64)
    codes64 = {k: v.double() for k, v in codes.items()}
    t0 = time.time()
    ref = run(orc64, img.double(), codes64, {r: ("f64", "f64") for r in ("trunk", "fpn", "head")})
    print(f"# fp64 reference: {time.time() - t0:.1f} s", flush=True)

    S = "f16"
    X = "f16x2"
    scenarios = [
        ("fp32 CPU oracle (what the parity tests compare with)", None),
        ("f16 everywhere (engine round 1, 1 MMA)", {"trunk": (S, S), "fpn": (S, S), "head": (S, S)}),
        ("weights f16x2, activations f16 (2 MMAs)", {"trunk": (S, X), "fpn": (S, X), "head": (S, X)}),
        ("activations f16x2, weights f16 (2 MMAs)", {"trunk": (X, S), "fpn": (X, S), "head": (X, S)}),
        ("f16x2 everywhere (3 MMAs)", {"trunk": (X, X), "fpn": (X, X), "head": (X, X)}),
        ("f16x2 trunk only", {"trunk": (X, X), "fpn": (S, S), "head": (S, S)}),
        ("f16x2 trunk + fpn", {"trunk": (X, X), "fpn": (X, X), "head": (S, S)}),
        ("f16x2 fpn + head only", {"trunk": (S, S), "fpn": (X, X), "head": (X, X)}),
        ("f16x2 head only", {"trunk": (S, S), "fpn": (S, S), "head": (X, X)}),
        ("bf16x2 everywhere (3 MMAs)", {r: ("bf16x2", "bf16x2") for r in ("trunk", "fpn", "head")}),
        ("tf32 (truncated) everywhere", {r: ("tf32t", "tf32t") for r in ("trunk", "fpn", "head")}),
    ]
    lines = []
    hdr = "| scenario | p3 | p5 | p7 | logits p3 | logits p5 | logits p7 | reg p3 | ctr p3 | ctr p3 (L2) | max score err |"
    lines.append(hdr)
    lines.append("|" + "---|" * 11)
    print(hdr, flush=True)
    for name, pol in scenarios:
        if args.only and args.only not in name:
            continue
        t0 = time.time()
        if pol is None:
            orc32 = MetaFCOSOracle(cfg, state)
            with torch.no_grad():
                il = orc32.preprocess([img])
                feats = orc32.features(il.tensor)
                logits, regs, ctrs, _ = orc32.head(feats, codes)
            got = ([f.double() for f in feats], [x.double() for x in logits], [x.double() for x in regs], [x.double() for x in ctrs])
        else:
            got = run(orc64, img.double(), codes64, pol)
        f, lg, rg, ct = got
        rf, rl, rr, rc = ref
        score = max(float(((lg[l].sigmoid() * ct[l].sigmoid()).sqrt() - (rl[l].sigmoid() * rc[l].sigmoid()).sqrt()).abs().max())
                    for l in range(5))
        row = (f"| {name} | {rel_err(f[0], rf[0]):.2e} | {rel_err(f[2], rf[2]):.2e} | {rel_err(f[4], rf[4]):.2e} | "
               f"{rel_err(lg[0], rl[0]):.2e} | {rel_err(lg[2], rl[2]):.2e} | {rel_err(lg[4], rl[4]):.2e} | "
               f"{rel_err(rg[0], rr[0]):.2e} | {rel_err(ct[0], rc[0]):.2e} | {rel_l2(ct[0], rc[0]):.2e} | {score:.2e} |")
        lines.append(row)
        print(row + f"   # {time.time() - t0:.0f} s", flush=True)
    if args.out:
        with open(args.out, "a") as fo:
            fo.write(f"\n### {args.h}x{args.w} {'smooth' if args.smooth else 'uniform-noise'} image, R-{args.depth}, seed {args.seed}\n\n")
            fo.write("max |a - b| / max |b| against the float64 oracle unless stated\n\n")
            fo.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
