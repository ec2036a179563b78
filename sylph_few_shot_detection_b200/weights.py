"""State-dict schema of `MetaOneStageDetector` (Meta-FCOS, CodeGenerator) and a deterministic synthetic fill.

The key layout is the boundary that lets reference checkpoints load unchanged:
  backbone.*                     upstream detectron2/AdelaiDet names (ResNet `bottom_up`, `fpn_lateral/output{3,4,5}`,
                                 `top_block.p6/p7`)
  proposal_generator.fcos_head.* `sylph/modeling/meta_fcos/fcos.py:72-122` (towers), `:422-440` (predictors, scales)
  code_generator.*               `sylph/modeling/code_generator/code_generator.py:328-333, 359-374, 509-581, 648-688`
  pixel_mean / pixel_std         `sylph/modeling/meta_arch/meta_one_stage_detector.py:60-65`

`synthetic_state_dict` gives every tensor a value that depends only on (key, seed): there are no datasets or
checkpoints offline, and the oracle, the golden generator (which fills the REFERENCE model with it) and the CUDA
path must all see identical weights regardless of module construction order.  Scales are chosen "trained-like"
(signal survives the towers, a few hundred candidates pass the 0.05 threshold) rather than the reference's
N(0, 0.01) initialisers, under which every logit equals the prior bias and detection is a no-op.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

_RESNET_BLOCKS = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}


def _conv_out_channels(cfg) -> int:
    return int(cfg.MODEL.META_LEARN.CODE_GENERATOR.OUT_CHANNEL)


def state_spec(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape, in module registration order."""
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def bn(prefix: str, c: int) -> None:
        for s in ("weight", "bias", "running_mean", "running_var"):
            spec[f"{prefix}.{s}"] = (c,)

    def conv(prefix: str, cout: int, cin: int, k: int, bias: bool) -> None:
        spec[f"{prefix}.weight"] = (cout, cin, k, k)
        if bias:
            spec[f"{prefix}.bias"] = (cout,)

    def gn(prefix: str, c: int) -> None:
        spec[f"{prefix}.weight"] = (c,)
        spec[f"{prefix}.bias"] = (c,)

    fpn_c = int(cfg.MODEL.FPN.OUT_CHANNELS)
    # FPN registers its own convs (reversed stage order is a Python-list detail; registration is res3, res4, res5)
    stage_channels = {3: 512, 4: 1024, 5: 2048}
    for stage in (3, 4, 5):
        conv(f"backbone.fpn_lateral{stage}", fpn_c, stage_channels[stage], 1, True)
        conv(f"backbone.fpn_output{stage}", fpn_c, fpn_c, 3, True)
    if int(cfg.MODEL.FCOS.TOP_LEVELS) == 2:
        conv("backbone.top_block.p6", fpn_c, fpn_c, 3, True)
        conv("backbone.top_block.p7", fpn_c, fpn_c, 3, True)
    conv("backbone.bottom_up.stem.conv1", 64, 3, 7, False)
    bn("backbone.bottom_up.stem.conv1.norm", 64)
    in_c, out_c, bott = 64, 256, 64
    for i, n in enumerate(_RESNET_BLOCKS[int(cfg.MODEL.RESNETS.DEPTH)]):
        for b in range(n):
            p = f"backbone.bottom_up.res{i + 2}.{b}"
            if in_c != out_c:
                conv(f"{p}.shortcut", out_c, in_c, 1, False)
                bn(f"{p}.shortcut.norm", out_c)
            conv(f"{p}.conv1", bott, in_c, 1, False)
            bn(f"{p}.conv1.norm", bott)
            conv(f"{p}.conv2", bott, bott, 3, False)
            bn(f"{p}.conv2.norm", bott)
            conv(f"{p}.conv3", out_c, bott, 1, False)
            bn(f"{p}.conv3.norm", out_c)
            in_c = out_c
        out_c *= 2
        bott *= 2

    head = "proposal_generator.fcos_head"
    for tower, n in (("cls_tower", int(cfg.MODEL.FCOS.NUM_CLS_CONVS)), ("bbox_tower", int(cfg.MODEL.FCOS.NUM_BOX_CONVS))):
        for i in range(n):
            conv(f"{head}.{tower}.{3 * i}", fpn_c, fpn_c, 3, True)
            gn(f"{head}.{tower}.{3 * i + 1}", fpn_c)
    k_logits = int(cfg.MODEL.FCOS.CLS_LOGITS_KERNEL_SIZE)
    conv(f"{head}.cls_logits", int(cfg.MODEL.FCOS.NUM_CLASSES), fpn_c, k_logits, True)
    conv(f"{head}.bbox_pred", 4, fpn_c, 3, True)
    conv(f"{head}.ctrness", 1, fpn_c, 3, True)
    conv(f"{head}.iou_overlap", 1, fpn_c, 3, True)
    if cfg.MODEL.FCOS.USE_SCALE:
        for lvl in range(len(cfg.MODEL.FCOS.IN_FEATURES)):
            spec[f"{head}.scales.{lvl}.scale"] = (1,)

    G = cfg.MODEL.META_LEARN.CODE_GENERATOR
    if not cfg.MODEL.META_LEARN.EPISODIC_LEARNING:
        pass                                            # base detector: no code generator (meta_one_stage_detector.py:83-87)
    elif G.NAME == "ROIEncoder":
        _roi_encoder_spec(cfg, spec, conv, gn)
        spec[f"{head}.cond_cls_logits.scales.0.scale"] = (1,)   # CondConvBlock(weight_len=256), head_utils.py:131-137
    else:
        _code_generator_spec(cfg, spec, conv, gn, fpn_c)
    spec["pixel_mean"] = (3, 1, 1)
    spec["pixel_std"] = (3, 1, 1)
    return spec


def _code_generator_spec(cfg, spec, conv, gn, fpn_c) -> None:
    """`CodeGenerator` / `CodeGeneratorHead` (code_generator.py:328-333, 359-374, 509-581, 648-688)."""
    G = cfg.MODEL.META_LEARN.CODE_GENERATOR
    cg = "code_generator.code_generator_head"
    oc = _conv_out_channels(cfg)
    for lvl in range(len(cfg.MODEL.FCOS.IN_FEATURES)):
        gn(f"{cg}.init_norm.{lvl}", fpn_c)
    idx = 0
    for norm_type, act in G.TOWER_LAYERS:
        conv(f"{cg}.support_set_shared_tower.{idx}", 256, 256, 3, True)
        idx += 1
        if norm_type in ("GN", "LN", "NaiveGN"):
            gn(f"{cg}.support_set_shared_tower.{idx}", 256)
            idx += 1
        if act in ("ReLU", "Tanh"):
            idx += 1
    if G.POST_NORM != "":
        gn(f"{cg}.post_norm", oc)
    if len(G.CLS_LAYER) == 3:
        conv(f"{cg}.support_set_cls_conv.0", oc, 256, 3, True)
        if G.CLS_LAYER[0] in ("GN", "LN", "NaiveGN"):
            gn(f"{cg}.support_set_cls_conv.1", oc)
    if len(G.BIAS_LAYER) == 3:
        conv(f"{cg}.support_set_cls_bias.0", 1, 256, 3, True)
        spec[f"{cg}.bias_scale.scale"] = (1,)
    if len(G.WEIGHT_LAYER) == 3:      # per-shot weight head (code_generator.py:583-612): conv 256 -> 1 (+ pool)
        conv(f"{cg}.support_set_cls_weight.0", 1, 256, 3, True)
    if G.USE_WEIGHT_SCALE and (G.CONV_L2_NORM or G.POST_NORM != ""):
        spec[f"{cg}.conv_scale.scale"] = (1,)


def _roi_encoder_spec(cfg, spec, conv, gn) -> None:
    """`ROIEncoder` (sylph/modeling/code_generator/roi_encoder.py:206-281): FeatureFusionModuleV2 with MS_CAM
    (utils.py:70-165), Tokenizer (:26-79), nn.TransformerEncoder, two HyperNetworkHeads (:82-115)."""
    G = cfg.MODEL.META_LEARN.CODE_GENERATOR
    cg = "code_generator"
    cam = f"{cg}.box_pooler.context_attention_module"
    for branch, (i1, i2, i3, i4) in (("local_att", (0, 1, 3, 4)), ("global_att", (1, 2, 4, 5))):
        conv(f"{cam}.{branch}.{i1}", 64, 256, 1, True)
        gn(f"{cam}.{branch}.{i2}", 64)
        conv(f"{cam}.{branch}.{i3}", 256, 64, 1, True)
        gn(f"{cam}.{branch}.{i4}", 256)
    conv(f"{cg}.box_pooler.conv.0", 256, 256, 3, True)
    gn(f"{cg}.box_pooler.conv.1", 256)
    T = G.TOKENIZER
    cin = 256
    for k in range(int(T.NUM_CONV)):
        conv(f"{cg}.tokenizer.conv{k + 1}", int(T.CONV_DIM), cin, 3, T.NORM == "")
        if T.NORM != "":
            gn(f"{cg}.tokenizer.conv{k + 1}.norm", int(T.CONV_DIM))
        cin = int(T.CONV_DIM)
    din = cin * int(G.ROI_BOX.POOLER_RESOLUTION) ** 2
    for k in range(int(T.NUM_FC)):
        spec[f"{cg}.tokenizer.fc{k + 1}.weight"] = (int(T.FC_DIM), din)
        spec[f"{cg}.tokenizer.fc{k + 1}.bias"] = (int(T.FC_DIM),)
        din = int(T.FC_DIM)
    d = int(T.FC_DIM)
    for l in range(int(G.TRANSFORMER_ENCODER.LAYERS)):
        p = f"{cg}.transformer_encoder.layers.{l}"
        spec[f"{p}.self_attn.in_proj_weight"] = (3 * d, d)
        spec[f"{p}.self_attn.in_proj_bias"] = (3 * d,)
        spec[f"{p}.self_attn.out_proj.weight"] = (d, d)
        spec[f"{p}.self_attn.out_proj.bias"] = (d,)
        spec[f"{p}.linear1.weight"] = (4 * d, d)
        spec[f"{p}.linear1.bias"] = (4 * d,)
        spec[f"{p}.linear2.weight"] = (d, 4 * d)
        spec[f"{p}.linear2.bias"] = (d,)
        for n in ("norm1", "norm2"):
            gn(f"{p}.{n}", d)
    for name, out_dim in (("weight_head", int(G.HEAD.OUTPUT_DIM)), ("bias_head", 1)):
        din = d
        for i in range(int(G.HEAD.NUM_FC)):
            dout = out_dim if i == int(G.HEAD.NUM_FC) - 1 else int(G.HEAD.FC_DIM)
            spec[f"{cg}.{name}.fc{i + 1}.weight"] = (dout, din)
            spec[f"{cg}.{name}.fc{i + 1}.bias"] = (dout,)
            din = dout

def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def synthetic_tensor(cfg, key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)

    def normal(std: float, mean: float = 0.0) -> torch.Tensor:
        return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

    def uniform(lo: float, hi: float) -> torch.Tensor:
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    if key == "pixel_mean":
        return torch.tensor(cfg.MODEL.PIXEL_MEAN, dtype=torch.float32).view(3, 1, 1)
    if key == "pixel_std":
        return torch.tensor(cfg.MODEL.PIXEL_STD, dtype=torch.float32).view(3, 1, 1)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith(".scale"):
        if "conv_scale" in key:
            return torch.full(shape, 6.0)
        if "bias_scale" in key:
            return torch.full(shape, 1.5)
        return uniform(0.8, 1.2)
    if leaf == "running_mean":
        return normal(0.1)
    if leaf == "running_var":
        return uniform(0.5, 1.5)
    if len(shape) == 1:  # norm affine or conv bias
        is_norm = (".norm." in key) or ("init_norm" in key) or ("post_norm" in key) or _is_gn_key(key)
        if leaf == "weight":
            if key.endswith("conv3.norm.weight"):
                return uniform(0.1, 0.4)  # damp the residual branches so features stay O(10) like a trained net
            if key.endswith("stem.conv1.norm.weight"):
                return uniform(0.02, 0.06)  # pixel-scale (0..255) inputs -> O(1) activations
            return uniform(0.5, 1.5)
        if is_norm:
            return normal(0.1)
        if key.endswith("cls_logits.bias"):
            return torch.full(shape, -math.log((1 - cfg.MODEL.FCOS.PRIOR_PROB) / cfg.MODEL.FCOS.PRIOR_PROB))
        if key.endswith("bbox_pred.bias"):
            return torch.full(shape, 1.0)
        return normal(0.02)
    if len(shape) == 2:  # nn.Linear / attention projections of the ROIEncoder generator
        if "bias_head" in key and shape[0] == 1:
            return normal(0.5 * math.sqrt(1.0 / shape[1]))
        if "weight_head" in key and shape[0] == int(cfg.MODEL.META_LEARN.CODE_GENERATOR.HEAD.OUTPUT_DIM) and shape[1] != 256:
            return normal(0.12 * math.sqrt(1.5 / shape[1]))   # keeps |logit| < ~10: no saturated-sigmoid score ties
        return normal(math.sqrt(1.5 / shape[1]))
    cout, cin, kh, kw = shape
    if key.startswith("backbone.bottom_up"):
        return normal(math.sqrt(2.0 / (cout * kh * kw)))
    if key.startswith("backbone."):
        bound = math.sqrt(3.0 / (cin * kh * kw))
        return uniform(-bound, bound)
    if key.endswith("cls_logits.weight"):
        return normal(0.01)
    if any(s in key for s in ("bbox_pred", "ctrness", "iou_overlap", "support_set_cls_bias")):
        return normal(0.03)
    if "support_set_cls_weight" in key:   # per-shot weight head: logits that spread the softmax weights well away from 1 / K
        return normal(0.3)
    return normal(math.sqrt(2.0 / (cin * kh * kw)))


def _is_gn_key(key: str) -> bool:
    """GroupNorm entries inside nn.Sequential towers sit at index 1 mod 3 (conv, GN, ReLU)."""
    parts = key.split(".")
    if len(parts) >= 2 and parts[-2].isdigit():
        owner = parts[-3] if len(parts) >= 3 else ""
        idx = int(parts[-2])
        if owner in ("cls_tower", "bbox_tower", "support_set_shared_tower"):
            return idx % 3 == 1
        if owner in ("support_set_cls_conv",):
            return idx == 1
        if owner == "conv" and "box_pooler" in key:
            return idx == 1
        if owner == "local_att":
            return idx in (1, 4)
        if owner == "global_att":
            return idx in (2, 5)
    return False


def synthetic_state_dict(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, synthetic_tensor(cfg, k, shp, seed)) for k, shp in state_spec(cfg).items())


def reference_init_state_dict(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """The reference's own initialisers ("random-init weights" of BASELINE.json): c2_msra for the ResNet,
    c2_xavier for FPN / P6P7, identity FrozenBN, N(0, 0.01) convolutions with zero bias in the FCOS head
    (sylph/modeling/meta_fcos/fcos.py:445-461) and in the code generator (code_generator.py:400-415), GroupNorm affine
    (1, 0), every `Scale` at its init value.  With these weights every class logit sits at the prior
    (sigmoid(-4.595) = 0.01 < INFERENCE_TH_TEST): detection has ZERO candidates -- the bench reports this variant next to
    the "trained-like" synthetic weights that exercise the proposal / NMS kernels."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    head_like = ("proposal_generator.fcos_head.", "code_generator.")
    for key, shape in state_spec(cfg).items():
        g = _gen(key, seed)
        leaf = key.rsplit(".", 1)[-1]
        if key in ("pixel_mean", "pixel_std"):
            t = synthetic_tensor(cfg, key, shape, seed)
        elif key.endswith(".scale"):
            t = torch.full(shape, 1.0)
        elif leaf == "running_mean":
            t = torch.zeros(shape)
        elif leaf == "running_var":
            t = torch.full(shape, 1.0 - 1e-5)           # detectron2 FrozenBatchNorm2d default
        elif len(shape) == 1:
            is_norm = (".norm." in key) or ("init_norm" in key) or ("post_norm" in key) or _is_gn_key(key) or "norm" in key.split(".")[-2]
            if key.endswith("cls_logits.bias"):
                t = torch.full(shape, -math.log((1 - cfg.MODEL.FCOS.PRIOR_PROB) / cfg.MODEL.FCOS.PRIOR_PROB))
            elif is_norm and leaf == "weight":
                t = torch.ones(shape)
            else:
                t = torch.zeros(shape)                  # norm bias, conv bias
        elif len(shape) == 2:
            bound = math.sqrt(1.0 / shape[1])           # nn.Linear default (kaiming_uniform, a = sqrt(5))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            cout, cin, kh, kw = shape
            if key.startswith(head_like):
                t = torch.randn(shape, generator=g) * 0.01
            elif key.startswith("backbone.bottom_up"):
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (cout * kh * kw))     # c2_msra_fill
            else:
                bound = math.sqrt(3.0 / (cin * kh * kw))                                    # c2_xavier_fill
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[key] = t.to(torch.float32)
    return out


def load_into_module(module: torch.nn.Module, state: Dict[str, torch.Tensor]) -> None:
    """Strict load: every key of the module must be present with the same shape (and vice versa)."""
    own = module.state_dict()
    missing = [k for k in own if k not in state]
    extra = [k for k in state if k not in own]
    if missing or extra:
        raise KeyError(f"state-dict mismatch: missing={missing[:5]} extra={extra[:5]}")
    for k, v in own.items():
        if tuple(v.shape) != tuple(state[k].shape):
            raise ValueError(f"shape mismatch for {k}: {tuple(v.shape)} vs {tuple(state[k].shape)}")
    module.load_state_dict(state, strict=True)
