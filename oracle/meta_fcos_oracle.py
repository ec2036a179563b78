"""ORACLE (test infrastructure only -- the product path never imports this; it must fail loudly without its CUDA
extension instead).

CPU fp32 restatement of the reference's Meta-FCOS few-shot INFERENCE path, written functionally over a flat
state dict (keys = SURVEY.md Appendix C).  Each function cites the reference lines it follows.  The un-vendored
upstream operators come from oracle/upstream.py.

PINNING: checked against outputs of the reference's own modules (run unmodified through oracle/shims by
oracle/make_golden.py) stored in tests/golden/*.pt -- see tests/test_oracle.py.  The reference's own tests hold no
numeric golden vectors for this path (shapes/keys only).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import upstream as up


def _gn(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, groups: int = 32) -> torch.Tensor:
    return F.group_norm(x, groups, w, b, 1e-5)


class MetaFCOSOracle:
    def __init__(self, cfg, state: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: v.detach().to(dtype) for k, v in state.items()}
        self.backbone = up.build_fcos_resnet_fpn_backbone(cfg).to(dtype)
        bb = {k[len("backbone."):]: v for k, v in self.sd.items() if k.startswith("backbone.")}
        self.backbone.load_state_dict(bb, strict=True)
        self.backbone.eval()
        self.in_features = list(cfg.MODEL.FCOS.IN_FEATURES)
        self.strides = list(cfg.MODEL.FCOS.FPN_STRIDES)
        G = cfg.MODEL.META_LEARN.CODE_GENERATOR
        self.G = G
        self.pooler = up.ROIPooler(output_size=G.ROI_BOX.POOLER_RESOLUTION, scales=[1.0 / s for s in self.strides],
                                   sampling_ratio=0, pooler_type=G.ROI_BOX.POOLER_TYPE)
        p = cfg.MODEL.FCOS.PRIOR_PROB
        self.bias_value = torch.tensor(-math.log((1 - p) / p))  # code_generator.py:422-423

    # -------------------------------------------------------------------------------- images / backbone
    def preprocess(self, images: Sequence[torch.Tensor]) -> up.ImageList:
        """meta_one_stage_detector.py:174-178: (x - mean) / std per image, then zero-pad to a /32 batch."""
        mean, std = self.sd["pixel_mean"], self.sd["pixel_std"]
        normed = [(im.to(self.dtype) - mean) / std for im in images]
        return up.ImageList.from_tensors(normed, self.backbone.size_divisibility)

    @torch.no_grad()
    def features(self, batch: torch.Tensor) -> List[torch.Tensor]:
        """meta_one_stage_detector.py:180-182, :247-250: backbone dict -> list ordered by FCOS.IN_FEATURES."""
        out = self.backbone(batch)
        return [out[f] for f in self.in_features]

    # -------------------------------------------------------------------------------- code generator
    def _cg(self, name: str) -> torch.Tensor:
        return self.sd["code_generator.code_generator_head." + name]

    def _cg_tower(self, x: torch.Tensor) -> torch.Tensor:
        """support_set_shared_tower (code_generator.py:648-688): per layer conv3x3 [+ norm] [+ act]."""
        idx = 0
        for norm_type, act in self.G.TOWER_LAYERS:
            x = F.conv2d(x, self._cg(f"support_set_shared_tower.{idx}.weight"),
                         self._cg(f"support_set_shared_tower.{idx}.bias"), padding=1)
            idx += 1
            if norm_type in ("GN", "NaiveGN", "LN"):
                groups = 1 if norm_type == "LN" else 32
                x = _gn(x, self._cg(f"support_set_shared_tower.{idx}.weight"),
                        self._cg(f"support_set_shared_tower.{idx}.bias"), groups)
                idx += 1
            elif norm_type not in ("", "none", None):
                raise NotImplementedError(norm_type)
            if act == "ReLU":
                x = F.relu(x)
                idx += 1
            elif act == "Tanh":
                x = torch.tanh(x)
                idx += 1
        return x

    @torch.no_grad()
    def roi_features(self, feats: List[torch.Tensor], boxes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """code_generator.py:928-930: one box per support image -> (N, 256, 7, 7); also returns level indices."""
        box_lists = [up.Boxes(boxes[i:i + 1].to(torch.float32)) for i in range(boxes.shape[0])]
        levels = up.assign_boxes_to_levels(box_lists, self.pooler.min_level, self.pooler.max_level,
                                           self.pooler.canonical_box_size, self.pooler.canonical_level)
        return self.pooler(feats, box_lists), levels

    @torch.no_grad()
    def per_shot_codes(self, roi: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """code_generator.py:941-967: shared tower, cls-conv + global average pool, bias conv (+ L2 over the 49
        positions when BIAS_L2_NORM) + pool.  Returns (N, OC, 1, 1) and (N, 1, 1, 1) or None."""
        x = self._cg_tower(roi) if len(self.G.TOWER_LAYERS) > 0 else roi
        w = F.conv2d(x, self._cg("support_set_cls_conv.0.weight"), self._cg("support_set_cls_conv.0.bias"), padding=1)
        if self.G.CLS_LAYER[0] in ("GN", "NaiveGN", "LN"):
            w = _gn(w, self._cg("support_set_cls_conv.1.weight"), self._cg("support_set_cls_conv.1.bias"),
                    1 if self.G.CLS_LAYER[0] == "LN" else 32)
        if self.G.CLS_LAYER[1] == "ReLU":
            w = F.relu(w)
        elif self.G.CLS_LAYER[1] == "Tanh":
            w = torch.tanh(w)
        k_s = int(self.G.CLS_LAYER[2])
        w = F.adaptive_avg_pool2d(w, (k_s, k_s))
        b = None
        if len(self.G.BIAS_LAYER) == 3:
            b = F.conv2d(x, self._cg("support_set_cls_bias.0.weight"), self._cg("support_set_cls_bias.0.bias"), padding=1)
            if self.G.BIAS_L2_NORM:
                shp = b.shape
                b = F.normalize(b.reshape(shp[0], shp[1], -1), p=2, dim=2).reshape(shp)
            b = F.adaptive_avg_pool2d(b, (1, 1))
        self._last_weight_logits = None
        if len(self.G.WEIGHT_LAYER) == 3:
            # support_set_cls_weight (code_generator.py:583-612, 969-973): conv 256 -> 1 (+ norm: none supported) + global pool
            wl = F.conv2d(x, self._cg("support_set_cls_weight.0.weight"), self._cg("support_set_cls_weight.0.bias"), padding=1)
            self._last_weight_logits = F.adaptive_avg_pool2d(wl, (1, 1))
        return w, b

    def shot_weights(self, n_cls: int, shot: int, dtype) -> torch.Tensor:
        """process_weight (code_generator.py:766-776) on the weight-head output of the LAST per_shot_codes call: softmax over the
        shots of every class; uniform 1 / K without a WEIGHT_LAYER (:805-806)."""
        wl = getattr(self, "_last_weight_logits", None)
        if wl is None:
            return torch.full((n_cls, shot, 1, 1, 1), 1.0 / shot, dtype=dtype)
        return torch.softmax(wl.view(n_cls, shot, 1, 1, 1), dim=1)

    @torch.no_grad()
    def class_code(self, support_images: Sequence[torch.Tensor], boxes: torch.Tensor) -> Dict[str, torch.Tensor]:
        """run_type="meta_learn_test_support" (meta_one_stage_detector.py:229-254 -> code_generator.py:924-1002).
        ALL rows of the call form ONE class (eval: num_shot = size(0), code_generator.py:790-793); uniform 1/K
        weights (:805-806, 817); codes are returned RAW (normalisation only `if self.training`, :993-994)."""
        il = self.preprocess(support_images)
        feats = self.features(il.tensor)
        roi, _ = self.roi_features(feats, boxes)
        w, b = self.per_shot_codes(roi)
        k = w.shape[0]
        weight = self.shot_weights(1, k, w.dtype)
        cls_conv = (weight * w.view(1, k, *w.shape[1:])).sum(dim=1)
        if b is not None:
            cls_bias = (weight * b.view(1, k, 1, 1, 1)).sum(dim=1)
        else:
            cls_bias = torch.zeros(1, 1, 1, 1, dtype=w.dtype)
        return {"cls_conv": cls_conv, "cls_bias": cls_bias}

    @torch.no_grad()
    def normalize_code(self, cls_conv: torch.Tensor, cls_bias: torch.Tensor,
                       cls_weight_norm: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """run_type="meta_learn_normalize_code": code_process_module (code_generator.py:864-875):
        post_norm GN (if present and C % 32 == 0, :838-839) -> L2 over channels (:841-842) -> x weight_norm ->
        x conv_scale (:873); bias: flatten -> x bias_scale -> + bias_value (:853-860)."""
        w = cls_conv
        assert w.ndim == 4
        if self.G.POST_NORM != "" and w.size(1) % 32 == 0:
            w = _gn(w, self._cg("post_norm.weight"), self._cg("post_norm.bias"), 1 if self.G.POST_NORM == "LN" else 32)
        if self.G.CONV_L2_NORM:
            w = F.normalize(w, p=2, dim=1)
        if cls_weight_norm is not None:
            w = w * cls_weight_norm
        key = "code_generator.code_generator_head.conv_scale.scale"
        if key in self.sd:
            w = w * self.sd[key]
        assert cls_bias.size(0) == 1, "predicted bias should only have batch size 1"
        b = cls_bias.reshape(cls_bias.numel())
        key = "code_generator.code_generator_head.bias_scale.scale"
        if key in self.sd:
            b = b * self.sd[key]
        b = b + self.bias_value.to(b.dtype)
        return w, b

    @staticmethod
    def pack_codes(codes: List[Dict]) -> Dict[str, torch.Tensor]:
        """format_class_codes_shared (sylph/evaluation/meta_learn_evaluation.py:71-103): order by
        support_set_target, concatenate, flatten the bias."""
        by_id = {int(c["support_set_target"]): c["class_code"] for c in codes}
        ids = sorted(by_id)
        return {"cls_conv": torch.cat([by_id[i]["cls_conv"] for i in ids], dim=0),
                "cls_bias": torch.cat([by_id[i]["cls_bias"].reshape(-1) for i in ids], dim=0)}

    # -------------------------------------------------------------------------------- FCOS head
    def _head(self, name: str) -> torch.Tensor:
        return self.sd["proposal_generator.fcos_head." + name]

    def _tower(self, x: torch.Tensor, which: str, n: int) -> torch.Tensor:
        """fcos.py:72-122: n x [conv3x3 + GN(32) + ReLU], weights shared across levels."""
        norm = self.cfg.MODEL.FCOS.NORM
        for i in range(n):
            x = F.conv2d(x, self._head(f"{which}.{3 * i}.weight"), self._head(f"{which}.{3 * i}.bias"), padding=1)
            if norm in ("GN", "NaiveGN"):
                x = _gn(x, self._head(f"{which}.{3 * i + 1}.weight"), self._head(f"{which}.{3 * i + 1}.bias"))
            elif norm not in ("none", None):
                raise NotImplementedError(norm)
            x = F.relu(x)
        return x

    @torch.no_grad()
    def head(self, feats: List[torch.Tensor], codes: Dict[str, torch.Tensor]):
        """MetaFCOSHead.forward (fcos.py:582-667) with CondConvBasic (head_utils.py:60-81)."""
        C = self.cfg.MODEL.FCOS
        k_s = int(self.G.CLS_LAYER[2])
        w = codes["cls_conv"].to(self.dtype)
        b = codes["cls_bias"].to(self.dtype) if self.G.USE_BIAS else None
        logits, regs, ctrs, ious = [], [], [], []
        for lvl, f in enumerate(feats):
            ct = self._tower(f, "cls_tower", int(C.NUM_CLS_CONVS))
            bt = self._tower(f, "bbox_tower", int(C.NUM_BOX_CONVS))
            logits.append(F.conv2d(ct, w, b, stride=1, padding=k_s // 2))
            reg = F.conv2d(bt, self._head("bbox_pred.weight"), self._head("bbox_pred.bias"), padding=1)
            if C.USE_SCALE:
                reg = reg * self._head(f"scales.{lvl}.scale")
            regs.append(F.relu(reg))
            ctrs.append(F.conv2d(bt, self._head("ctrness.weight"), self._head("ctrness.bias"), padding=1))
            ious.append(F.conv2d(bt, self._head("iou_overlap.weight"), self._head("iou_overlap.bias"), padding=1))
        return logits, regs, ctrs, ious

    @torch.no_grad()
    def level_candidates(self, level: int, logits: torch.Tensor, reg: torch.Tensor, ctr: torch.Tensor,
                         iou: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
        """forward_for_single_feature_map (fcos_outputs.py:904-1008), per image of the batch.  Candidates are
        returned in (location, class) order, i.e. the order of `nonzero()` (:967-969)."""
        C = self.cfg.MODEL.FCOS
        N, NC, H, W = logits.shape
        stride = self.strides[level]
        locations = up.compute_locations(H, W, stride, logits.device).to(self.dtype)
        cls = logits.permute(0, 2, 3, 1).reshape(N, -1, NC).sigmoid()
        box_reg = (reg * stride).permute(0, 2, 3, 1).reshape(N, -1, 4)  # fcos_outputs.py:786
        ctrness = ctr.permute(0, 2, 3, 1).reshape(N, -1).sigmoid()
        iouness = iou.permute(0, 2, 3, 1).reshape(N, -1).sigmoid()
        quality = sorted(C.BOX_QUALITY)
        if quality == ["ctrness"]:
            q = ctrness
        elif quality == ["iou"]:
            q = iouness
        elif quality == ["ctrness", "iou"]:
            q = torch.sqrt(iouness * ctrness)
        else:
            raise NotImplementedError(quality)
        if C.THRESH_WITH_CTR:
            cls = cls * q[:, :, None]
        cand = cls > C.INFERENCE_TH_TEST
        if not C.THRESH_WITH_CTR:
            cls = cls * q[:, :, None]
        out = []
        for i in range(N):
            nz = cand[i].nonzero()
            loc_idx, cls_idx = nz[:, 0], nz[:, 1]
            score = cls[i][cand[i]]
            top_n = min(int(cand[i].sum()), int(C.PRE_NMS_TOPK_TEST))
            if score.numel() > top_n:
                # torch.topk(sorted=False) leaves ties implementation-defined; the oracle keeps the smallest
                # (location, class) index among equal scores and tests treat exact ties at the cut as a guard band.
                order = torch.sort(score, descending=True, stable=True).indices[:top_n]
                order = torch.sort(order).values
                score, loc_idx, cls_idx = score[order], loc_idx[order], cls_idx[order]
            r = box_reg[i][loc_idx]
            l = locations[loc_idx]
            boxes = torch.stack([l[:, 0] - r[:, 0], l[:, 1] - r[:, 1], l[:, 0] + r[:, 2], l[:, 1] + r[:, 3]], dim=1)
            out.append({"boxes": boxes, "scores": torch.sqrt(score), "classes": cls_idx, "locations": l,
                        "loc_index": loc_idx, "levels": torch.full_like(cls_idx, level)})
        return out

    @torch.no_grad()
    def select(self, per_image: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """select_over_all_levels (fcos_outputs.py:1010-1028): class-aware NMS, then keep everything scoring at
        least the POST_NMS_TOPK-th best (ties kept, `>=`)."""
        C = self.cfg.MODEL.FCOS
        keep = up.batched_nms(per_image["boxes"].float(), per_image["scores"].float(), per_image["classes"], C.NMS_TH)
        res = {k: v[keep] for k, v in per_image.items()}
        n = keep.numel()
        topk = int(C.POST_NMS_TOPK_TEST)
        if n > topk > 0:
            thresh, _ = torch.kthvalue(res["scores"], n - topk + 1)
            m = res["scores"] >= thresh
            res = {k: v[m] for k, v in res.items()}
        return res

    @staticmethod
    def postprocess(res: Dict[str, torch.Tensor], image_size: Tuple[int, int], out_h: int, out_w: int):
        """detector_postprocess (meta_one_stage_detector.py:288-295, upstream A.8): rescale, clip, drop empty."""
        sx, sy = out_w / image_size[1], out_h / image_size[0]
        b = res["boxes"].clone()
        b[:, 0::2] *= sx
        b[:, 1::2] *= sy
        b[:, 0::2] = b[:, 0::2].clamp(min=0, max=out_w)
        b[:, 1::2] = b[:, 1::2].clamp(min=0, max=out_h)
        ok = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
        out = {k: v[ok] for k, v in res.items()}
        out["boxes"] = b[ok]
        return out

    @torch.no_grad()
    def detect(self, query_images: Sequence[torch.Tensor], codes: Dict[str, torch.Tensor],
               out_sizes: Optional[Sequence[Tuple[int, int]]] = None, return_intermediate: bool = False):
        """run_type="meta_learn_test_instance" (meta_one_stage_detector.py:261-296 -> fcos.py:184-268 ->
        fcos_outputs.py:743-812)."""
        il = self.preprocess(query_images)
        feats = self.features(il.tensor)
        logits, regs, ctrs, ious = self.head(feats, codes)
        per_level = [self.level_candidates(l, logits[l], regs[l], ctrs[l], ious[l]) for l in range(len(feats))]
        results = []
        pre_nms = []
        for i in range(len(query_images)):
            merged = {k: torch.cat([per_level[l][i][k] for l in range(len(feats))], dim=0) for k in per_level[0][i]}
            pre_nms.append(merged)
            sel = self.select(merged)
            h, w = out_sizes[i] if out_sizes is not None else il.image_sizes[i]
            results.append(self.postprocess(sel, il.image_sizes[i], h, w))
        if return_intermediate:
            return results, {"features": feats, "logits": logits, "reg": regs, "ctr": ctrs, "iou": ious,
                             "pre_nms": pre_nms, "image_sizes": il.image_sizes}
        return results

    # -------------------------------------------------------------------------------- training forward
    INF = 100000000          # fcos_outputs.py:29
    BACKGROUND_ID = 100000   # fcos_outputs.py:101

    @staticmethod
    def filter_gt(query_records: Sequence[Dict], support_targets: Sequence[int]):
        """MetaProposalNetwork._get_gt (meta_one_stage_detector.py:184-221): keep the ground truths of a query image
        whose class is one of the episode's support classes, in their original order."""
        keep_ids = [int(t) for t in support_targets]
        out = []
        for rec in query_records:
            boxes = rec["instances"].gt_boxes.tensor
            classes = rec["instances"].gt_classes
            sel = [i for i in range(len(classes)) if int(classes[i]) in keep_ids]
            out.append((boxes[sel].reshape(-1, 4).to(torch.float32), classes[sel].reshape(-1).to(torch.int64)))
        return out

    def fcos_targets(self, level_sizes: Sequence[Tuple[int, int]], gts):
        """FCOSOutputs._get_ground_truth / compute_targets_for_locations / get_sample_region
        (fcos_outputs.py:140-349).  Returns level-first flattened (L, N, H, W order, :125-138) labels (int64),
        target_inds (int64), reg_targets (fp32, divided by the level stride, :186-189) and the per-location level."""
        C = self.cfg.MODEL.FCOS
        soi = [-1] + list(C.SIZES_OF_INTEREST) + [self.INF]
        locs, ranges, num_loc = [], [], []
        for l, (h, w) in enumerate(level_sizes):
            loc = up.compute_locations(h, w, self.strides[l], torch.device("cpu")).to(torch.float32)
            locs.append(loc)
            num_loc.append(loc.shape[0])
            ranges.append(torch.tensor([soi[l], soi[l + 1]], dtype=torch.float32)[None].expand(loc.shape[0], 2))
        locations, ranges = torch.cat(locs), torch.cat(ranges)
        xs, ys = locations[:, 0], locations[:, 1]
        K = locations.shape[0]
        labels, regs, inds = [], [], []
        num_targets = 0
        for boxes, classes in gts:
            if boxes.numel() == 0:                                       # :274-281
                labels.append(torch.full((K,), self.BACKGROUND_ID, dtype=torch.int64))
                regs.append(torch.zeros(K, 4))
                inds.append(torch.full((K,), -1, dtype=torch.int64))
                continue
            area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
            l = xs[:, None] - boxes[:, 0][None]
            t = ys[:, None] - boxes[:, 1][None]
            r = boxes[:, 2][None] - xs[:, None]
            b = boxes[:, 3][None] - ys[:, None]
            reg = torch.stack([l, t, r, b], dim=2)
            if C.CENTER_SAMPLE:                                          # get_sample_region, :193-248
                cx = (boxes[:, 0] + boxes[:, 2]) * 0.5
                cy = (boxes[:, 1] + boxes[:, 3]) * 0.5
                if float(cx[0]) * K == 0:                                # ":213" quirk: first centre at x == 0
                    inside = torch.zeros(K, boxes.shape[0], dtype=torch.bool)
                else:
                    rad = torch.cat([torch.full((n,), float(self.strides[i] * C.POS_RADIUS))
                                     for i, n in enumerate(num_loc)])[:, None]
                    x0 = torch.maximum(cx[None] - rad, boxes[:, 0][None])
                    y0 = torch.maximum(cy[None] - rad, boxes[:, 1][None])
                    x1 = torch.minimum(cx[None] + rad, boxes[:, 2][None])
                    y1 = torch.minimum(cy[None] + rad, boxes[:, 3][None])
                    inside = torch.stack([xs[:, None] - x0, ys[:, None] - y0, x1 - xs[:, None], y1 - ys[:, None]],
                                         -1).min(-1)[0] > 0
            else:
                inside = reg.min(dim=2)[0] > 0
            mx = reg.max(dim=2)[0]
            cared = (mx >= ranges[:, [0]]) & (mx <= ranges[:, [1]])
            a = area[None].repeat(K, 1)
            a[~inside] = self.INF
            a[~cared] = self.INF
            amin, gi = a.min(dim=1)                                      # first minimum on ties
            lab = classes[gi].clone()
            lab[amin == self.INF] = self.BACKGROUND_ID
            labels.append(lab)
            regs.append(reg[torch.arange(K), gi])
            inds.append(gi + num_targets)
            num_targets += boxes.shape[0]

        def level_first(per_image):
            split = [torch.split(x, num_loc, dim=0) for x in per_image]
            return [torch.cat(lv, dim=0) for lv in zip(*split)]
        lab_l, reg_l, ind_l = level_first(labels), level_first(regs), level_first(inds)
        reg_l = [r / float(self.strides[i]) for i, r in enumerate(reg_l)]
        lvl = torch.cat([torch.full((x.shape[0],), i, dtype=torch.int64) for i, x in enumerate(lab_l)])
        return torch.cat(lab_l), torch.cat(ind_l), torch.cat(reg_l), lvl

    @staticmethod
    def iou_terms(pred: torch.Tensor, target: torch.Tensor):
        """IOULoss.compute_ious (sylph/modeling/meta_fcos/iou_loss.py:26-65) on (l, t, r, b) pairs."""
        ta = (target[:, 0] + target[:, 2]) * (target[:, 1] + target[:, 3])
        pa = (pred[:, 0] + pred[:, 2]) * (pred[:, 1] + pred[:, 3])
        wi = torch.min(pred[:, 0], target[:, 0]) + torch.min(pred[:, 2], target[:, 2])
        hi = torch.min(pred[:, 3], target[:, 3]) + torch.min(pred[:, 1], target[:, 1])
        gw = torch.max(pred[:, 0], target[:, 0]) + torch.max(pred[:, 2], target[:, 2])
        gh = torch.max(pred[:, 3], target[:, 3]) + torch.max(pred[:, 1], target[:, 1])
        ac = gw * gh
        inter = wi * hi
        union = ta + pa - inter
        ious = (inter + 1.0) / (union + 1.0)
        return ious, ious - (ac - union) / ac

    def fcos_losses(self, logits, regs, ctrs, gts, support_targets: Sequence[int], world_size: int = 1,
                    reduce=None) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        """FCOSOutputs.losses -> fcos_losses_episodic_learning (fcos_outputs.py:351-637) for box_on == False.
        `reduce` stands for adet's reduce_sum (identity for one process)."""
        C = self.cfg.MODEL.FCOS
        reduce = reduce or (lambda t: t)
        labels, gt_inds, reg_targets, lvl = self.fcos_targets([tuple(x.shape[-2:]) for x in logits], gts)
        flat = lambda xs, c: torch.cat([x.permute(0, 2, 3, 1).reshape(-1, c) for x in xs], dim=0)
        NC = logits[0].shape[1]
        logits_pred = flat(logits, NC)
        reg_pred = flat(regs, 4)
        ctr_pred = flat(ctrs, 1).reshape(-1)
        st = torch.tensor([int(t) for t in support_targets], dtype=torch.int64).view(1, -1)
        pos = torch.nonzero(labels != self.BACKGROUND_ID).squeeze(1)
        total_pos = float(reduce(torch.tensor([pos.numel()], dtype=torch.int64)).item())
        num_pos_avg = max(total_pos / world_size, 1.0)
        class_target = (st == labels[:, None]).to(torch.float32)
        loss_cls = up.sigmoid_focal_loss(logits_pred, class_target, alpha=C.LOSS_ALPHA, gamma=C.LOSS_GAMMA,
                                         reduction="sum") / num_pos_avg
        rt = reg_targets[pos]
        lr, tb = rt[:, [0, 2]], rt[:, [1, 3]]
        if rt.shape[0]:
            ctr_t = torch.sqrt((lr.min(dim=-1)[0] / lr.max(dim=-1)[0]) * (tb.min(dim=-1)[0] / tb.max(dim=-1)[0]))
        else:
            ctr_t = rt.new_zeros(0)
        denorm = max(float(reduce(ctr_t.sum()).item()) / world_size, 1e-6)
        ious, gious = self.iou_terms(reg_pred[pos], rt)
        if pos.numel() > 0:
            kind = C.LOC_LOSS_TYPE
            per = -torch.log(ious) if kind == "iou" else (1 - ious) if kind == "linear_iou" else (1 - gious)
            loss_loc = (per * ctr_t).sum() / denorm
            loss_ctr = F.binary_cross_entropy_with_logits(ctr_pred[pos], ctr_t, reduction="sum") / num_pos_avg
        else:
            loss_loc = reg_pred[pos].sum() * 0
            loss_ctr = ctr_pred[pos].sum() * 0
        losses = {"loss_fcos_cls": loss_cls}
        P = self.cfg.MODEL.PROPOSAL_GENERATOR
        if not (P.FREEZE_BBOX_BRANCH or P.FREEZE):                       # box_branch_loss_on, fcos_outputs.py:87-92
            losses.update({"loss_fcos_loc": loss_loc, "loss_fcos_ctr": loss_ctr})
        extras = {"labels": labels, "target_inds": gt_inds, "reg_targets": reg_targets, "fpn_levels": lvl,
                  "num_pos": torch.tensor(pos.numel()), "ctr_targets_sum": ctr_t.sum(), "loss_denorm": torch.tensor(denorm)}
        return losses, extras

    @torch.no_grad()
    def training_forward(self, batched_inputs: Sequence[Dict], world_size: int = 1, reduce=None):
        """forward_few_shot_detector_training (meta_one_stage_detector.py:325-388): one item per class with
        "support_set" (SHOT records, one selected box each), "query_set" and "support_set_target"."""
        return self._training_forward(batched_inputs, world_size, reduce)

    def _training_forward(self, batched_inputs: Sequence[Dict], world_size: int = 1, reduce=None):
        """Body of training_forward without the no_grad guard (training_grads differentiates it).  The backbone runs under
        no_grad either way: MODEL.BACKBONE.FREEZE is on in every shipped meta-training config."""
        cls = type(self)
        shot = int(self.cfg.MODEL.META_LEARN.SHOT)
        support = [r for x in batched_inputs for r in x["support_set"]]
        targets = [int(x["support_set_target"]) for x in batched_inputs]
        query = [r for x in batched_inputs for r in x["query_set"]]
        assert len(support) % shot == 0, f"Total size {len(support)} must be divisible by number of shot {shot}"
        with torch.no_grad():
            q_il = self.preprocess([r["image"] for r in query])
            q_feats = self.features(q_il.tensor)
            s_il = self.preprocess([r["image"] for r in support])
            s_feats = self.features(s_il.tensor)
        gts = self.filter_gt(query, targets)
        boxes = torch.stack([r["instances"].gt_boxes.tensor[0] for r in support])   # one GT per support image
        roi, _ = self.roi_features(s_feats, boxes)
        w, b = cls.per_shot_codes.__wrapped__(self, roi)
        n_cls = w.shape[0] // shot
        weight = self.shot_weights(n_cls, shot, w.dtype)                            # code_generator.py:766-776, 805-817
        raw_conv = (weight * w.view(n_cls, shot, *w.shape[1:])).sum(dim=1)
        raw_bias = (weight * b.view(n_cls, shot, 1, 1, 1)).sum(dim=1) if b is not None else torch.zeros(n_cls, 1, 1, 1)
        cls_conv, cls_bias = self.process_codes_training(raw_conv, raw_bias)        # :993-994
        codes = {"cls_conv": cls_conv, "cls_bias": cls_bias}
        logits, regs, ctrs, ious = cls.head.__wrapped__(self, q_feats, codes)
        losses, extras = self.fcos_losses(logits, regs, ctrs, gts, targets, world_size, reduce)
        extras.update({"codes": codes, "gts": gts, "raw_codes": {"cls_conv": raw_conv, "cls_bias": raw_bias}, "roi": roi})
        return losses, extras

    # parameters of the code generator that receive a gradient in the reference's training step (the `init_norm.*` layers
    # are registered but never called on this path)
    def trainable_code_generator_keys(self) -> List[str]:
        pre = "code_generator.code_generator_head."
        return [k for k, v in self.sd.items() if k.startswith(pre) and torch.is_floating_point(v)
                and not k.startswith(pre + "init_norm.")]

    def trainable_keys(self) -> List[str]:
        """Code generator + (unless FREEZE_CLS_TOWER / FREEZE) the FCOS class tower: the tensors the B200 path has backward
        kernels for; the reference additionally trains whatever else the configuration leaves unfrozen."""
        keys = self.trainable_code_generator_keys()
        P = self.cfg.MODEL.PROPOSAL_GENERATOR
        if not (P.FREEZE_CLS_TOWER or P.FREEZE):
            keys += [k for k in self.sd if k.startswith("proposal_generator.fcos_head.cls_tower.")]
        return keys

    def training_grads(self, batched_inputs: Sequence[Dict], world_size: int = 1, reduce=None):
        """Backward of the episodic training step for the CODE GENERATOR and the FCOS CLASS TOWER: autograd through
        the restated forward.  d(sum of the returned losses) / d(parameter) for every tensor of `trainable_keys`, as the reference's
        `losses = model(batched); sum(losses.values()).backward()` leaves them in `.grad` (detectron2 SimpleTrainer.run_step),
        plus the gradient with respect to the final class codes.  Returns (losses, {state_dict key: grad}, extras)."""
        keys = self.trainable_keys()
        saved = {k: self.sd[k] for k in keys}
        leaves = {k: saved[k].detach().clone().requires_grad_(True) for k in keys}
        self.sd.update(leaves)
        try:
            with torch.enable_grad():
                losses, extras = self._training_forward(batched_inputs, world_size, reduce)
                for t in (extras["codes"]["cls_conv"], extras["codes"]["cls_bias"]):
                    t.retain_grad()
                total = sum(losses.values())
                total.backward()
            grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])).detach() for k in keys}
            extras["grad_codes"] = {"cls_conv": extras["codes"]["cls_conv"].grad.detach(),
                                    "cls_bias": extras["codes"]["cls_bias"].grad.detach()}
        finally:
            self.sd.update(saved)
        return {k: v.detach() for k, v in losses.items()}, grads, extras

    def process_codes_training(self, cls_conv: torch.Tensor, cls_bias: torch.Tensor):
        """code_process_module on the whole (C, 256, 1, 1) batch (code_generator.py:864-875; the `size(0) == 1`
        assert of process_bias, :850-851, only applies in eval mode)."""
        w = cls_conv
        if self.G.POST_NORM != "" and w.size(1) % 32 == 0:
            w = _gn(w, self._cg("post_norm.weight"), self._cg("post_norm.bias"), 1 if self.G.POST_NORM == "LN" else 32)
        if self.G.CONV_L2_NORM:
            w = F.normalize(w, p=2, dim=1)
        key = "code_generator.code_generator_head.conv_scale.scale"
        if key in self.sd:
            w = w * self.sd[key]
        b = cls_bias.reshape(cls_bias.numel())
        key = "code_generator.code_generator_head.bias_scale.scale"
        if key in self.sd:
            b = b * self.sd[key]
        return w, b + self.bias_value.to(b.dtype)
