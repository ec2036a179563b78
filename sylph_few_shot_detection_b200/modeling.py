"""Host-side mirror of the reference's plugin surface for the Meta-FCOS inference path.

Same registry names, class names, `run_type` strings, argument meaning, return schema and error behaviour as
  * `CODE_GENERATOR_REGISTRY` / `build_code_generator` .. sylph/modeling/code_generator/build.py:18-39
  * `CodeGenerator.forward` ........................... sylph/modeling/code_generator/code_generator.py:1037-1053
  * `MetaFCOS` (PROPOSAL_GENERATOR_REGISTRY) .......... sylph/modeling/meta_fcos/fcos.py:158-268
  * `build_fcos_resnet_fpn_backbone` (BACKBONE_REGISTRY) configs/COCO-Detection/Meta-FCOS/Base-FCOS.yaml:3-4
  * `MetaOneStageDetector.forward(run_type=...)` ...... sylph/modeling/meta_arch/meta_one_stage_detector.py:425-445
so the reference's configs and runner loops (sylph/evaluation/meta_learn_evaluation.py:256-470) drive it unchanged.
All arithmetic happens in libsylph_b200.so; these classes hold no parameters and only marshal arguments.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .runtime import CODE_STRIDE, SLOT_QUERY, SLOT_SUPPORT, Engine
from .structures import HAVE_DETECTRON2, Boxes, Instances, Registry, ShapeSpec
from .weights import state_spec

CODE_GENERATOR_REGISTRY = Registry("CODE_GENERATOR")
CODE_GENERATOR_REGISTRY.__doc__ = "Registry for code generator (same name as the reference's)."

if HAVE_DETECTRON2:  # pragma: no cover - bind into the real registries when they exist
    from detectron2.modeling import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY  # type: ignore
else:
    META_ARCH_REGISTRY = Registry("META_ARCH")
    PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
    BACKBONE_REGISTRY = Registry("BACKBONE")


def build_code_generator(cfg, feature_channels: int, feature_levels: Optional[int], strides: Tuple[int]):
    """reference: sylph/modeling/code_generator/build.py:29-39 (None when the name is empty)."""
    name = cfg.MODEL.META_LEARN.CODE_GENERATOR.NAME
    if name == "":
        return None
    return CODE_GENERATOR_REGISTRY.get(name)(cfg, feature_channels, feature_levels, strides)


def select_a_mask(gt_instances: Sequence[Any], use_all_masks: bool = False) -> List[torch.Tensor]:
    """One GT box per support image, drawn with the global NumPy RNG exactly like the reference
    (sylph/modeling/code_generator/utils.py:27-47); empty boxes raise ValueError.  `use_all_masks`
    (CODE_GENERATOR.ALL_MASK): every box of the image, no RNG draw."""
    out = []
    for inst in gt_instances:
        boxes = inst.gt_boxes.tensor
        if len(boxes) == 0:
            raise ValueError("support image without a ground-truth box")
        if use_all_masks:
            out.append(boxes)
        else:
            idx = np.random.choice(range(len(boxes)), 1)
            out.append(boxes[idx])
    return out


def support_boxes(gt_instances: Sequence[Any], use_all_masks: bool, total_shots: int) -> torch.Tensor:
    """The (total_shots, 4) boxes the ROI pooler receives (code_generator.py:928-937): `select_a_mask`, then the reference's
    own check that pooling returned exactly one ROI per support image -- which, with ALL_MASK, holds only when every image
    carries a single box."""
    boxes = torch.cat([b.reshape(-1, 4).cpu() for b in select_a_mask(gt_instances, use_all_masks)], dim=0)
    assert boxes.shape[0] == total_shots, \
        f"pooled_features.shape[0] {boxes.shape[0]} Vs batch_size * num_shots {total_shots}"
    return boxes


class _EngineBound(nn.Module):
    """Mixin: modules on this path share the detector's C-ABI context."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_engine_ref", None)

    def bind_engine(self, engine: Engine) -> None:
        object.__setattr__(self, "_engine_ref", engine)

    @property
    def engine(self) -> Engine:
        eng = object.__getattribute__(self, "_engine_ref")
        if eng is None:
            raise RuntimeError("no weights loaded: call MetaOneStageDetector.load_state_dict first "
                               "(the B200 path has no CPU fallback)")
        return eng


@CODE_GENERATOR_REGISTRY.register()
class CodeGenerator(_EngineBound):
    """Drop-in for the reference's `CodeGenerator` (registered under the same name, same forward signature)."""

    def __init__(self, cfg, feature_channels: int, feature_levels: int, strides: Tuple[int]):
        super().__init__()
        if feature_channels != 256 or feature_levels != 5:   # the reference accepts any (build.py:29-39)
            raise NotImplementedError("the B200 path is built for the 256-channel p3..p7 pyramid of the shipped configs")
        self.in_features = cfg.MODEL.FCOS.IN_FEATURES
        self.strides = tuple(strides)
        self.all_mask = cfg.MODEL.META_LEARN.CODE_GENERATOR.ALL_MASK
        self.contrastive_loss = cfg.MODEL.META_LEARN.CODE_GENERATOR.CONTRASTIVE_LOSS

    def forward(self, features: Optional[List[torch.Tensor]], target_instances=None, cls_norm: bool = False,
                class_codes: Optional[List[Dict]] = None):
        if not self.training and cls_norm and class_codes is not None:
            return self.forward_normalize_code(class_codes)
        return self.forward_roi_align(features, target_instances)

    # generate mode: List[(N, 256, H_l, W_l)] NCHW features + List[Instances] -> raw code of ONE class (eval quirk:
    # num_shot = size(0), code_generator.py:790-793)
    def forward_roi_align(self, features: List[torch.Tensor], gt_instances: Sequence[Any]) -> Dict[str, torch.Tensor]:
        assert not self.training, "the B200 path implements inference only"
        total_shots = features[0].size(0)
        assert len(gt_instances) == total_shots
        boxes = support_boxes(gt_instances, self.all_mask, total_shots)
        h, w = features[0].shape[-2:]
        self.engine.import_features(SLOT_SUPPORT, features, (h * self.strides[0], w * self.strides[0]))
        return self._codes_of_one_class(boxes, list(range(total_shots)))

    def _codes_of_one_class(self, boxes: torch.Tensor, roi_image: List[int]) -> Dict[str, torch.Tensor]:
        raw = self.engine.generate_codes(SLOT_SUPPORT, boxes, roi_image, [0, len(roi_image)])
        out = {"cls_conv": raw[:, :256].reshape(1, 256, 1, 1), "cls_bias": raw[:, 256:].reshape(1, 1, 1, 1)}
        if self.contrastive_loss == "snnl":
            raise NotImplementedError("CONTRASTIVE_LOSS='snnl' output is not produced on the inference path")
        return out

    # normalise mode (code_generator.py:877-897): mutates the dicts in place, returns the same list
    def forward_normalize_code(self, codes: List[Dict]) -> List[Dict]:
        assert not self.training, "Only support testing mode here"
        assert codes is not None
        if len(codes) == 0:
            return codes
        convs, biases = [], []
        dev = self.engine.device  # codes normally already live on the device: no host round trip, no sync
        for code in codes:
            assert "class_code" in code, "class_code is not in code"
            assert "cls_conv" in code["class_code"], "class_conv is not in class_code"
            if "cls_weight_norm" in code["class_code"]:
                raise NotImplementedError("cls_weight_norm (SCALE_LAYER) is not supported")
            b = code["class_code"]["cls_bias"]
            assert b.numel() == 1, "predicted bias should only have batch size 1"
            convs.append(code["class_code"]["cls_conv"].reshape(1, 256).to(dev, torch.float32))
            biases.append(b.reshape(1, 1).to(dev, torch.float32))
        # three concatenations whatever the number of classes (was two small kernels per class)
        raw = torch.cat([torch.cat(convs, dim=0), torch.cat(biases, dim=0)], dim=1)
        normed = self.engine.normalize_codes(raw)
        for i, code in enumerate(codes):
            code["class_code"]["cls_conv"] = normed[i, :256].reshape(1, 256, 1, 1)
            code["class_code"]["cls_bias"] = normed[i, 256:].reshape(1)
        return codes


@CODE_GENERATOR_REGISTRY.register()
class ROIEncoder(_EngineBound):
    """Drop-in for the reference's `ROIEncoder` generator (sylph/modeling/code_generator/roi_encoder.py:118-204), same
    registry name and forward signature: ROIAlign -> FeatureFusionModuleV2 (conv + MS_CAM context gate) -> Tokenizer ->
    TransformerEncoder -> class-token mean -> weight / bias hyper-network heads, all inside `sylph_generate_codes`.

    Returns {"cls_conv": (bs, 256, 1, 1), "cls_bias": (bs,)}; the bias already contains the focal-loss prior, so the
    codes are final.  Reference quirk: `normalize_class_code` passes `features=None, cls_norm=True, class_codes=...`
    (meta_one_stage_detector.py:256-259), keywords `ROIEncoder.forward` does not accept -- the reference meta-test loop
    raises TypeError there.  This class accepts them and returns the codes unchanged (there is nothing to normalise)."""

    def __init__(self, cfg, feature_channels: int, feature_levels: int, strides: Tuple[int]):
        super().__init__()
        if feature_channels != 256 or feature_levels != 5:   # the reference accepts any (build.py:29-39)
            raise NotImplementedError("the B200 path is built for the 256-channel p3..p7 pyramid of the shipped configs")
        self.cfg = cfg
        self.strides = tuple(strides)
        self.shot = cfg.MODEL.META_LEARN.SHOT
        self.eval_shot = cfg.MODEL.META_LEARN.EVAL_SHOT
        self.bias_value = -float(np.log((1 - 0.01) / 0.01))   # roi_encoder.py:141-142

    def forward(self, support_set_image_features: Optional[List[torch.Tensor]] = None, gt_instances=None, *,
                features: Optional[List[torch.Tensor]] = None, target_instances=None, cls_norm: bool = False,
                class_codes: Optional[List[Dict]] = None):
        if cls_norm and class_codes is not None:
            return class_codes
        feats = support_set_image_features if support_set_image_features is not None else features
        insts = gt_instances if gt_instances is not None else target_instances
        assert not self.training, "the B200 path implements inference only"
        num_shots = self.eval_shot
        boxes = torch.cat([b.reshape(-1, 4)[:1] for b in select_a_mask(insts)], dim=0)
        n = feats[0].shape[0]
        assert n % num_shots == 0, f"{n} % {num_shots}"                       # roi_encoder.py:160-163
        for lvl, f in enumerate(feats):
            assert f.shape[0] == n, f"lvl {lvl}, {f.shape[0]} vs {n}"
        assert len(insts) == n, f"boxes_ls {len(insts)} vs {n}"
        h, w = feats[0].shape[-2:]
        self.engine.import_features(SLOT_SUPPORT, feats, (h * self.strides[0], w * self.strides[0]))
        raw = self.engine.generate_codes(SLOT_SUPPORT, boxes, list(range(n)), list(range(0, n + 1, num_shots)))
        return {"cls_conv": raw[:, :256].reshape(-1, 256, 1, 1), "cls_bias": raw[:, 256].reshape(-1)}


def _register_alias(registry, name: str, obj) -> None:
    if hasattr(registry, "_do_register"):  # fvcore / detectron2 Registry
        registry._do_register(name, obj)
    else:
        registry._map[name] = obj


_register_alias(CODE_GENERATOR_REGISTRY, "CodeGeneratorHead", CodeGenerator)  # the reference registers both names


class _FCOSHeadHandle(nn.Module):
    """Attribute tree the reference's freezing code touches (meta_one_stage_detector.py:117-155); no parameters."""

    def __init__(self):
        super().__init__()
        for name in ("cls_tower", "bbox_tower", "share_tower", "cls_logits", "bbox_pred", "ctrness", "iou_overlap"):
            self.add_module(name, nn.Sequential())


@PROPOSAL_GENERATOR_REGISTRY.register()
class MetaFCOS(_EngineBound):
    """Drop-in for `MetaFCOS` (fcos.py:158-268): eval forward returns (List[Instances], {})."""

    def __init__(self, cfg, input_shape: Dict[str, Any]):
        super().__init__()
        self.cfg = cfg
        self.in_features = cfg.MODEL.FCOS.IN_FEATURES
        self.fpn_strides = cfg.MODEL.FCOS.FPN_STRIDES
        self.fcos_head = _FCOSHeadHandle()
        self.in_channels_to_top_module = 256

    def forward(self, images, features, gt_instances=None, top_module=None, support_set_per_class_code=None,
                support_set_targets=None):
        if self.training:
            # training branch of MetaFCOS.forward (fcos.py:217-246): (results, losses); no proposals are yielded
            assert support_set_per_class_code is not None and support_set_targets is not None, \
                "only the episodic (code-conditioned) losses are implemented on the B200 path"
            if features is not None:
                feats = [features[f] for f in self.in_features]
                h, w = feats[0].shape[-2:]
                self.engine.import_features(SLOT_QUERY, feats, (h * self.fpn_strides[0], w * self.fpn_strides[0]))
            return {}, self.losses(support_set_per_class_code, support_set_targets, gt_instances)
        if support_set_per_class_code is None:
            raise NotImplementedError("base-detector inference (cls_logits) is not implemented on the B200 path")
        image_sizes = images.image_sizes if hasattr(images, "image_sizes") else images
        if features is not None:  # NCHW features from a foreign backbone
            feats = [features[f] for f in self.in_features]
            h, w = feats[0].shape[-2:]
            # the slot learns the true image sizes: with out_sizes == image_sizes the fused postprocess scales by exactly 1
            # (the reference's proposal generator returns un-scaled boxes in the frame of the padded batch)
            self.engine.import_features(SLOT_QUERY, feats, (h * self.fpn_strides[0], w * self.fpn_strides[0]),
                                        image_sizes=[tuple(int(v) for v in s) for s in image_sizes])
        return self.predict(support_set_per_class_code, image_sizes, image_sizes), {}

    def predict_device(self, class_codes: Dict[str, torch.Tensor], out_sizes, code_rows: Optional[torch.Tensor] = None,
                       codes_ready=None):
        """Head + proposals + NMS + postprocess, everything left on the device and nothing synchronised:
        (dets (n, max_dets, 9) fp32, counts (n,) int32) -- rows in descending score order, see SYLPH_DET_STRIDE.
        `code_rows` + `codes_ready`: (C, 257) rows packed on ANOTHER stream and the event recorded there."""
        if code_rows is not None:
            return self.engine.detect(SLOT_QUERY, code_rows, out_sizes, codes_ready=codes_ready)
        codes = pack_code_rows(class_codes).to(self.engine.device)
        return self.engine.detect(SLOT_QUERY, codes, out_sizes)

    def predict(self, class_codes: Dict[str, torch.Tensor], image_sizes, out_sizes) -> List[Any]:
        dets, counts = self.predict_device(class_codes, out_sizes)
        counts = counts.cpu().tolist()      # synchronises: the overflow word of this call is on the host as well
        self.engine.detect_poll()
        return instances_from_detections(dets, counts, out_sizes)


def instances_from_detections(dets: torch.Tensor, counts: Sequence[int], out_sizes) -> List[Any]:
    """(n, max_dets, 9) detection rows -> the reference's `Instances` fields (fcos_outputs.py:986-1006)."""
    results = []
    for i, n in enumerate(counts):
        d = dets[i, :n]
        inst = Instances(tuple(int(v) for v in out_sizes[i]))
        inst.pred_boxes = Boxes(d[:, 0:4])
        inst.scores = d[:, 4]
        inst.pred_classes = d[:, 5].to(torch.int64)
        inst.locations = d[:, 6:8]
        inst.fpn_levels = d[:, 8].to(torch.int64)
        results.append(inst)
    return results


def _box_branch_loss_on(cfg) -> bool:
    P = cfg.MODEL.PROPOSAL_GENERATOR
    return not (P.FREEZE_BBOX_BRANCH or P.FREEZE)                       # fcos_outputs.py:87-92


def _reduce_sum(t: torch.Tensor) -> torch.Tensor:
    """adet.utils.comm.reduce_sum (fcos_outputs.py:522,558): all-reduce SUM, identity for one process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return t
    t = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def _world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _meta_fcos_losses(self, class_codes: Dict[str, torch.Tensor], support_set_targets: Sequence[Any],
                      gt_instances: Sequence[Any], want_targets: bool = False):
    """FCOSOutputs.losses -> fcos_losses_episodic_learning (fcos_outputs.py:351-637) on the features in SLOT_QUERY:
    one fused kernel assigns every location its ground truth and accumulates the loss sums; the two `reduce_sum`
    scalars (positives, centre-ness target sum) travel as ONE 2-element all-reduce of a device tensor."""
    assert support_set_targets is not None
    assert gt_instances is not None and len(gt_instances) > 0
    eng = self.engine
    codes = pack_code_rows(class_codes).to(eng.device)
    targets = [int(t) for t in support_set_targets]
    boxes = [g.gt_boxes.tensor.reshape(-1, 4).cpu().float() for g in gt_instances]
    classes = [g.gt_classes.reshape(-1).cpu().to(torch.int64) for g in gt_instances]
    offsets = [0]
    for b in boxes:
        offsets.append(offsets[-1] + b.shape[0])
    box_on = _box_branch_loss_on(self.cfg)
    if getattr(eng, "_loss_box_branch", True) != box_on:      # box losses off (FREEZE_BBOX_BRANCH): the head skips the box branch
        eng.set_loss_box_branch(box_on)
        eng._loss_box_branch = box_on
    res = eng.fcos_loss_sums(SLOT_QUERY, codes, targets, torch.cat(boxes), torch.cat(classes), offsets, want_targets)
    sums, extra = res if want_targets else (res, None)
    world = _world_size()
    glob = _reduce_sum(sums[1:3]) if world > 1 else None
    out = eng.fcos_loss_finalize(sums, glob, world)
    losses = {"loss_fcos_cls": out[0]}
    if _box_branch_loss_on(self.cfg):
        losses.update({"loss_fcos_loc": out[1], "loss_fcos_ctr": out[2]})
    return (losses, {"sums": sums, "labels": extra[0], "target_inds": extra[1], "reg_targets": extra[2],
                     "global_pos_ctr": glob}) if want_targets else losses


MetaFCOS.losses = _meta_fcos_losses


def pack_code_rows(class_codes: Dict[str, torch.Tensor]) -> torch.Tensor:
    """{"cls_conv": (C, 256, 1, 1), "cls_bias": (C,)} -> (C, 257) rows of the C ABI."""
    w = class_codes["cls_conv"]
    assert w.dim() == 4 and w.shape[1] == 256 and w.shape[2] == 1 and w.shape[3] == 1, f"cls_conv has shape {tuple(w.shape)}"
    b = class_codes["cls_bias"].reshape(-1, 1)
    assert b.shape[0] == w.shape[0]
    return torch.cat([w.reshape(w.shape[0], 256).float(), b.float().to(w.device)], dim=1).contiguous()


class _BackboneHandle(_EngineBound):
    """`build_fcos_resnet_fpn_backbone` drop-in: `size_divisibility`, `output_shape()` and `forward(x)` on a normalised,
    zero-padded (N, 3, H, W) batch -> {"p3".."p7": (N, 256, H_l, W_l) fp32}, the call the reference makes at
    meta_one_stage_detector.py:180-182.  (MetaOneStageDetector itself hands RAW images to the engine: the
    normalisation is fused into the stem's input preparation.)"""

    size_divisibility = 32

    def __init__(self, cfg):
        super().__init__()
        self._out = {f"p{3 + i}": ShapeSpec(channels=cfg.MODEL.FPN.OUT_CHANNELS, stride=8 << i) for i in range(5)}

    def output_shape(self):
        return self._out

    def forward(self, x):
        self.engine.extract_features_normalized(SLOT_SUPPORT, x)
        return {f"p{3 + l}": self.engine.export_features(SLOT_SUPPORT, l) for l in range(5)}


@BACKBONE_REGISTRY.register()
def build_fcos_resnet_fpn_backbone(cfg, input_shape=None):
    return _BackboneHandle(cfg)


def build_backbone(cfg):
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, None)


def build_proposal_generator(cfg, input_shape):
    return PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, input_shape)


class _CodeGeneratorGrad(torch.autograd.Function):
    """Ties the loss value the engine computed to the code generator's parameters: backward runs the engine's backward
    kernels (sylph_fcos_cls_loss_backward + sylph_codegen_backward) and hands every parameter its gradient, so that
    `sum(losses.values()).backward()` fills `.grad` like the reference's autograd graph does."""

    @staticmethod
    def forward(ctx, loss_value, closure, *params):
        ctx.closure = closure
        return loss_value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None) + tuple(ctx.closure(grad_out))


def _register_parameter_tree(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    """Register `param` under its dotted state_dict name below `root` (plain container modules for the inner names),
    so that named_parameters() yields the reference's keys."""
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        if name not in m._modules:
            m.add_module(name, nn.Module())
        m = m._modules[name]
    m.register_parameter(parts[-1], param)


@META_ARCH_REGISTRY.register()
class MetaOneStageDetector(nn.Module):
    """Drop-in for the reference meta-architecture in eval mode (meta_one_stage_detector.py:415-455)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.episodic_learning = cfg.MODEL.META_LEARN.EPISODIC_LEARNING
        self.backbone = build_backbone(cfg)
        self.proposal_generator = build_proposal_generator(cfg, self.backbone.output_shape())
        shapes = [self.backbone.output_shape()[f] for f in cfg.MODEL.FCOS.IN_FEATURES]
        self.code_generator = (build_code_generator(cfg, feature_channels=shapes[0].channels,
                                                    feature_levels=len(shapes), strides=cfg.MODEL.FCOS.FPN_STRIDES)
                               if self.episodic_learning else None)
        if self.episodic_learning:
            assert self.code_generator is not None
        self.register_buffer("pixel_mean", torch.Tensor(cfg.MODEL.PIXEL_MEAN).view(-1, 1, 1))
        self.register_buffer("pixel_std", torch.Tensor(cfg.MODEL.PIXEL_STD).view(-1, 1, 1))
        self.in_features = cfg.MODEL.FCOS.IN_FEATURES
        self._state: Dict[str, torch.Tensor] = {}
        self._engine: Optional[Engine] = None
        self._trainable: Dict[str, nn.Parameter] = {}      # code-generator parameters (enable_code_generator_training)
        self._synced_versions: Tuple[int, ...] = ()
        self._train_step = 0
        # operand precision of the engine built by load_state_dict: None = $SYLPH_PRECISION, else "exact"
        # (runtime.Engine); assign "fast" before loading the weights to trade the fp32-level agreement for speed
        self.precision: Optional[str] = None
        self.eval()

    # ------------------------------------------------------------------ weights: reference key layout (Appendix C)
    def load_state_dict(self, state_dict, strict: bool = True):  # type: ignore[override]
        spec = state_spec(self.cfg)
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:4]} unexpected {unexpected[:4]}")
        for k, shp in spec.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {tuple(shp)}")
        self._state = {k: state_dict[k].detach().cpu().float() for k in spec if k in state_dict}
        dev = self.pixel_mean.device
        index = dev.index if dev.type == "cuda" and dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._engine = Engine(self.cfg, index, self.precision)
        self._engine.load_state_dict(self._state)
        for m in (self.backbone, self.proposal_generator, self.code_generator):
            if m is not None:
                m.bind_engine(self._engine)
        if self._trainable:                                  # reloaded weights: the parameters follow
            with torch.no_grad():
                for k, p in self._trainable.items():
                    p.copy_(self._state[k].to(p.device))
            self._synced_versions = tuple(p._version for p in self._trainable.values())
            self._engine.set_training(self.training and self.trains_cls_tower)   # a fresh engine: re-apply the training mode
        elif self.training:
            self.enable_code_generator_training()
        return self

    def state_dict(self, *args, **kwargs):  # type: ignore[override]
        out = dict(self._state)
        for k, p in self._trainable.items():        # the live values of the parameters an optimiser may have stepped
            out[k] = p.detach().cpu().float().clone()
        return out

    # ------------------------------------------------------------------ trainable code generator (meta-training)
    def train(self, mode: bool = True):  # type: ignore[override]
        super().train(mode)
        if mode and self._engine is not None:
            self.enable_code_generator_training()
        if self._engine is not None and self._trainable:
            # the head keeps the class tower's activations only while training (eval passes stay on the two ping-pong buffers)
            self._engine.set_training(bool(mode) and self.trains_cls_tower)
        return self

    def enable_code_generator_training(self) -> List[nn.Parameter]:
        """Register every `code_generator.*` tensor as an nn.Parameter on the device under its reference name (the
        reference's CodeGenerator owns them as module parameters, code_generator.py:520-646), so that an optimiser built
        from `model.parameters()` trains the hyper-network.  Called by `train()` once weights are loaded.  The detector
        itself stays frozen on this path: there are no backward kernels for the backbone or the FCOS towers, which is the
        shipped hyper-network stage (configs/COCO-Detection/Meta-FCOS/Meta-FCOS-finetune-lvis.yaml)."""
        if self._trainable or not self.episodic_learning or isinstance(self.code_generator, ROIEncoder):
            return list(self._trainable.values())
        if self.cfg.MODEL.META_LEARN.CODE_GENERATOR.FREEZE:
            return []
        P = self.cfg.MODEL.PROPOSAL_GENERATOR
        if not (self.cfg.MODEL.BACKBONE.FREEZE and P.FREEZE_BBOX_BRANCH) and not P.FREEZE:
            import logging
            logging.getLogger(__name__).warning(
                "the configuration leaves the backbone / the box branch trainable (BACKBONE.FREEZE / FREEZE_BBOX_BRANCH); the B200 path "
                "has backward kernels for the code generator and the FCOS class tower only -- those parts stay frozen")
        dev = self.engine.device
        for k, v in self._state.items():
            if k.startswith("code_generator.") and torch.is_floating_point(v):
                p = nn.Parameter(v.to(dev, torch.float32).contiguous().clone())
                _register_parameter_tree(self.code_generator, k[len("code_generator."):], p)
                self._trainable[k] = p
        if self.trains_cls_tower:
            pre = "proposal_generator.fcos_head.cls_tower."
            for k, v in self._state.items():
                if k.startswith(pre):
                    p = nn.Parameter(v.to(dev, torch.float32).contiguous().clone())
                    _register_parameter_tree(self.proposal_generator, k[len("proposal_generator."):], p)
                    self._trainable[k] = p
            self.engine.set_training(self.training)
        self._synced_versions = tuple(p._version for p in self._trainable.values())
        return list(self._trainable.values())

    @property
    def trains_cls_tower(self) -> bool:
        """The class tower trains with the code generator unless PROPOSAL_GENERATOR.FREEZE_CLS_TOWER / FREEZE (the reference freezes
        by `requires_grad = False`, meta_one_stage_detector.py:117-155)."""
        P = self.cfg.MODEL.PROPOSAL_GENERATOR
        return not (P.FREEZE_CLS_TOWER or P.FREEZE)

    def _sync_code_generator(self) -> None:
        """Hand the engine the parameters' current values when an optimiser (or anything else) has changed them in place."""
        if not self._trainable:
            return
        versions = tuple(p._version for p in self._trainable.values())
        if versions != self._synced_versions:
            # device-side refresh (sylph_update_code_generator_device / sylph_update_cls_tower_device); state_dict() reads the
            # live parameters
            live = {k: p.detach() for k, p in self._trainable.items()}
            self.engine.update_code_generator_device(live)
            if self.trains_cls_tower:
                self.engine.update_cls_tower_device(live)
            self._synced_versions = versions

    @property
    def device(self):
        return self._engine.device if self._engine is not None else self.pixel_mean.device

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise RuntimeError("no weights loaded: call load_state_dict first (the B200 path has no CPU fallback)")
        return self._engine

    # ------------------------------------------------------------------ run_type protocol
    def forward(self, batched_inputs, class_code=None, run_type=None):
        self._sync_code_generator()      # parameters an optimiser has stepped since the last call reach the engine first
        if run_type is None and self.training:
            # MetaProposalNetwork.forward (meta_one_stage_detector.py:388-412): losses of one training episode batch
            if not self.episodic_learning:
                raise NotImplementedError("base-detector pre-training forward is not implemented on the B200 path")
            return self.forward_few_shot_detector_training(batched_inputs)
        if self.training:
            raise NotImplementedError(f"not support this forward type: {run_type}, class_code: {class_code}")
        if run_type is None:
            # "a normal base detector inference" (meta_one_stage_detector.py:435-441 -> MetaProposalNetwork.forward :388-412)
            processed_results = self.forward_base_detector(batched_inputs)
            renamed = [{"instances": r["proposals"]} for r in processed_results]
            return processed_results if len(renamed) == 0 else renamed
        if run_type == "meta_learn_test_support":
            return self.forward_class_code(batched_inputs)
        if run_type == "meta_learn_normalize_code":
            return self.normalize_class_code(class_code)
        if run_type == "meta_learn_test_instance":
            return self.forward_instances(batched_inputs, class_code)
        raise NotImplementedError(f"not support this forward type: {run_type}, class_code: {class_code}")

    def forward_base_detector(self, batched_inputs: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
        """meta_one_stage_detector.py:298-323 in eval mode: backbone, FCOS head with the model's own `cls_logits` classifier
        (MetaFCOSHead.forward_base_train, fcos.py:544-576), proposals, NMS, detector_postprocess -> [{"proposals": Instances}].
        On the engine the `cls_logits` weights are simply the class-code rows of the code-conditioned classifier."""
        if self.episodic_learning:   # MetaProposalNetwork.forward raises exactly this for an episodic model in eval mode
            raise NotImplementedError("Episodic learning inferrence for image and features is not supported in forward.")
        assert not self.training
        k = int(self.cfg.MODEL.FCOS.CLS_LOGITS_KERNEL_SIZE)
        if k != 1:
            raise NotImplementedError("the B200 classifier is the 1x1 convolution of the shipped configs (CLS_LOGITS_KERNEL_SIZE 1)")
        codes = {"cls_conv": self._state["proposal_generator.fcos_head.cls_logits.weight"],
                 "cls_bias": self._state["proposal_generator.fcos_head.cls_logits.bias"]}
        return [{"proposals": r["instances"]} for r in self._detect(batched_inputs, codes)]

    def forward_class_code(self, batched_inputs: List[Dict[str, Any]]) -> Dict[str, torch.Tensor]:
        """One class per call (assert at meta_one_stage_detector.py:238)."""
        assert not self.training, "Not for training"
        assert len(batched_inputs) == 1, f"batched_inputs has length: {len(batched_inputs)}"
        return self.forward_class_codes_batched(batched_inputs)[0]

    # images of one trunk pass of the batched support path, in units of 800 x 1344 images (activation memory bound)
    SUPPORT_PASS_IMAGE_BUDGET = 64

    @staticmethod
    def class_padded_size(item: Dict[str, Any], divisibility: int = 32) -> Tuple[int, int]:
        """The size ImageList.from_tensors pads ONE reference call (= one class, meta_one_stage_detector.py:174-178) to."""
        hs = [int(r["image"].shape[-2]) for r in item["support_set"]]
        ws = [int(r["image"].shape[-1]) for r in item["support_set"]]
        return (-(-max(hs) // divisibility) * divisibility, -(-max(ws) // divisibility) * divisibility)

    def forward_class_codes_batched(self, batched_inputs: List[Dict[str, Any]],
                                    features_in_slot: bool = False) -> List[Dict[str, torch.Tensor]]:
        """B200-native extension: the support sets of MANY classes through shared backbone batches and one
        code-generation launch sequence per batch; result[i] equals forward_class_code([batched_inputs[i]]).
        The reference runs one class per call and pads each call to ITS OWN maximum size, so only classes that pad to the
        same size share a trunk pass (features near the right / bottom border and p6 / p7 depend on the padded size), and
        a pass holds at most SUPPORT_PASS_IMAGE_BUDGET full-size images (LVIS-scale class lists do not fit one batch).
        `features_in_slot=True`: the support pyramid of exactly these images already sits in SLOT_SUPPORT
        (Engine.extract_features_multi; the caller has checked that all classes pad to one size)."""
        roi_encoder = isinstance(self.code_generator, ROIEncoder)
        if roi_encoder:
            for item in batched_inputs:   # bs = 1 per call in the reference: N % EVAL_SHOT == 0
                n = len(item["support_set"])
                assert n % self.code_generator.eval_shot == 0 and n == self.code_generator.eval_shot, \
                    f"{n} % {self.code_generator.eval_shot}"
        # one host-RNG draw per support image, in the order of the reference's per-class calls (utils.py:27-47)
        all_mask = bool(self.cfg.MODEL.META_LEARN.CODE_GENERATOR.ALL_MASK) and not roi_encoder
        boxes_of = [support_boxes([r["instances"] for r in item["support_set"]], all_mask, len(item["support_set"]))
                    for item in batched_inputs]
        if features_in_slot:
            passes = [list(range(len(batched_inputs)))]
        else:
            by_size: Dict[Tuple[int, int], List[int]] = {}
            for i, item in enumerate(batched_inputs):
                by_size.setdefault(self.class_padded_size(item), []).append(i)
            passes = []
            for (hp, wp), idx in by_size.items():
                per_image = (hp * wp) / float(800 * 1344)
                cur, load = [], 0.0
                for i in idx:
                    cost = len(batched_inputs[i]["support_set"]) * per_image
                    if cur and load + cost > self.SUPPORT_PASS_IMAGE_BUDGET:
                        passes.append(cur)
                        cur, load = [], 0.0
                    cur.append(i)
                    load += cost
                if cur:
                    passes.append(cur)
        rows: List[Optional[torch.Tensor]] = [None] * len(batched_inputs)
        for idx in passes:
            records, offsets = [], [0]
            for i in idx:
                records.extend(batched_inputs[i]["support_set"])
                offsets.append(len(records))
            if not features_in_slot:
                self.engine.extract_features(SLOT_SUPPORT, [r["image"] for r in records])
            raw = self.engine.generate_codes(SLOT_SUPPORT, torch.cat([boxes_of[i] for i in idx], dim=0),
                                             list(range(len(records))), offsets)
            for k, i in enumerate(idx):
                rows[i] = raw[k:k + 1]
        if roi_encoder:
            return [{"cls_conv": r[:, :256].reshape(1, 256, 1, 1), "cls_bias": r[0, 256:].reshape(1)} for r in rows]
        return [{"cls_conv": r[:, :256].reshape(1, 256, 1, 1), "cls_bias": r[:, 256:].reshape(1, 1, 1, 1)} for r in rows]

    # ------------------------------------------------------------------ training forward + code-generator backward
    def _get_gt(self, batched_inputs: List[Dict[str, Any]], support_set_targets=None):
        """meta_one_stage_detector.py:184-221: keep the ground truths whose class is one of the episode's classes."""
        if "instances" not in batched_inputs[0]:
            if "targets" in batched_inputs[0]:
                raise NotImplementedError("targets are not supported")
            return None
        if support_set_targets is None:
            return [x["instances"] for x in batched_inputs]
        assert isinstance(support_set_targets, list), "support_set_targets is not list"
        wanted = [int(t) for t in support_set_targets]
        out = []
        for x in batched_inputs:
            boxes = x["instances"].gt_boxes.tensor.reshape(-1, 4)
            classes = x["instances"].gt_classes.reshape(-1)
            keep = [i for i in range(len(classes)) if int(classes[i]) in wanted]
            inst = Instances(tuple(x["instances"].image_size))
            inst.gt_boxes = Boxes(boxes[keep].reshape(-1, 4).to(torch.float32))
            inst.gt_classes = classes[keep].to(torch.int64)
            out.append(inst)
        return out

    def forward_few_shot_detector_training(self, batched_inputs: List[Dict[str, Any]], want_targets: bool = False):
        """meta_one_stage_detector.py:325-388: each item = the query + support set of one class.  Returns the loss dict
        (`loss_fcos_cls` [, `loss_fcos_loc`, `loss_fcos_ctr`]) as device scalars.  With the code generator's parameters
        registered (`model.train()` after load_state_dict) and autograd enabled, `loss_fcos_cls` carries a backward hook:
        `sum(losses.values()).backward()` fills `.grad` of every `code_generator.*` parameter like the reference's
        autograd graph (the detector is frozen on this path)."""
        assert self.training
        assert "support_set" in batched_inputs[0]
        assert "query_set" in batched_inputs[0]
        if isinstance(self.code_generator, ROIEncoder):
            raise NotImplementedError("ROIEncoder training forward (transformer dropout) is not implemented")
        support = [r for x in batched_inputs for r in x["support_set"]]
        targets = [x["support_set_target"] for x in batched_inputs]
        query = [r for x in batched_inputs for r in x["query_set"]]
        shot = int(self.cfg.MODEL.META_LEARN.SHOT)
        assert len(support) % shot == 0, \
            f"Total size {len(support)} must be divisible by number of shot {shot}"     # code_generator.py:787-789
        gts = self._get_gt(query, support_set_targets=targets)
        boxes = torch.cat([b.reshape(-1, 4)[:1].cpu() for b in select_a_mask([r["instances"] for r in support])], dim=0)
        self.engine.extract_features_multi([(SLOT_QUERY, [r["image"] for r in query]),
                                            (SLOT_SUPPORT, [r["image"] for r in support])])
        self._sync_code_generator()
        class_offsets = list(range(0, len(support) + 1, shot))
        raw = self.engine.generate_codes(SLOT_SUPPORT, boxes, list(range(len(support))), class_offsets)
        codes = self.engine.normalize_codes(raw)                                        # code_generator.py:993-994
        class_codes = {"cls_conv": codes[:, :256].reshape(-1, 256, 1, 1), "cls_bias": codes[:, 256].reshape(-1)}
        differentiable = bool(self._trainable) and torch.is_grad_enabled()
        res = self.proposal_generator.losses(class_codes, targets, gts, want_targets or differentiable)
        if not differentiable:
            return res
        losses, extra = res
        # ---- backward hook: loss_fcos_cls is the only loss that depends on the code generator (the box losses read the
        # frozen box branch alone); its gradient runs through the engine's backward kernels
        self._train_step += 1
        step, eng = self._train_step, self.engine
        keys = [k for k, p in self._trainable.items() if p.requires_grad]
        params = [self._trainable[k] for k in keys]
        live = {k: self._trainable[k].detach() for k in self._trainable}
        tgt = [int(t) for t in targets]
        world = _world_size()
        glob = extra["global_pos_ctr"]      # {positives, centre-ness target sum} over all ranks (None for one process)

        def closure(grad_out):
            if self._train_step != step:
                raise RuntimeError("backward() of a training episode must run before the next forward: the engine's activation "
                                   "buffers of that episode have been overwritten")
            g_codes = eng.fcos_cls_loss_backward(SLOT_QUERY, len(tgt), tgt, extra["labels"], extra["sums"], glob, world, grad_out)
            grads = eng.codegen_backward(class_offsets, raw, g_codes, live)
            if self.trains_cls_tower:
                grads.update(eng.cls_tower_backward(SLOT_QUERY, codes, tgt, extra["labels"], extra["sums"], live, glob, world, grad_out))
            self._last_grad_codes, self._last_final_codes = g_codes, codes
            return [grads[k] if k in grads and not k.startswith("code_generator.code_generator_head.init_norm.") else None
                    for k in keys]

        losses = dict(losses)
        losses["loss_fcos_cls"] = _CodeGeneratorGrad.apply(losses["loss_fcos_cls"], closure, *params)
        return (losses, extra) if want_targets else losses

    def normalize_class_code(self, codes: List[Dict]):
        assert self.episodic_learning
        assert not self.training
        return self.code_generator(features=None, target_instances=None, cls_norm=True, class_codes=codes)

    def forward_instances(self, batched_inputs: List[Dict[str, Any]], class_codes: Dict[str, torch.Tensor],
                          features_in_slot: bool = False):
        assert self.episodic_learning
        assert not self.training, "Not for training"
        return self._detect(batched_inputs, class_codes, features_in_slot)

    def _detect(self, batched_inputs: List[Dict[str, Any]], class_codes: Dict[str, torch.Tensor], features_in_slot: bool = False):
        images = [x["image"] for x in batched_inputs]
        if not features_in_slot:
            self.engine.extract_features(SLOT_QUERY, images)
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
        out_sizes = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)]
        results = self.proposal_generator.predict(class_codes, sizes, out_sizes)
        return [{"instances": r} for r in results]

    def forward_instances_device(self, batched_inputs: List[Dict[str, Any]], class_codes: Dict[str, torch.Tensor],
                                 features_in_slot: bool = False, code_rows: Optional[torch.Tensor] = None, codes_ready=None):
        """`forward_instances` without the final device-to-host synchronisation: returns (dets, counts, out_sizes) with
        the detections still on the device (`instances_from_detections` turns them into Instances).  Lets a caller
        enqueue the next episode before it reads this one's results (runner.EpisodePipeline.run_async)."""
        assert self.episodic_learning
        assert not self.training, "Not for training"
        images = [x["image"] for x in batched_inputs]
        if not features_in_slot:
            self.engine.extract_features(SLOT_QUERY, images)
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
        out_sizes = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)]
        dets, counts = self.proposal_generator.predict_device(class_codes, out_sizes, code_rows, codes_ready)
        return dets, counts, out_sizes


def build_model(cfg, precision: Optional[str] = None) -> MetaOneStageDetector:
    """`runner.build_model(cfg)` equivalent: META_ARCH_REGISTRY lookup by cfg.MODEL.META_ARCHITECTURE.
    `precision` ("exact" | "fast" | None) is handed to the engine when the weights are loaded."""
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.precision = precision
    return model
