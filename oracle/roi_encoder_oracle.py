"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the reference's ROIEncoder code generator and of the CondConvBlock classifier it selects in
the FCOS head, on top of MetaFCOSOracle (backbone, ROI pooling, towers, proposals):
  * ROIEncoder.forward ............. sylph/modeling/code_generator/roi_encoder.py:146-204
  * FeatureFusionModuleV2.forward .. sylph/modeling/code_generator/utils.py:144-165  (+ GlobalAdaptiveAvgPool2d :51-67)
  * MS_CAM.forward ................. sylph/modeling/code_generator/utils.py:70-103
  * Tokenizer / HyperNetworkHead ... roi_encoder.py:26-115
  * CondConvBlock.forward .......... sylph/modeling/meta_fcos/head_utils.py:121-162 (selected at fcos.py:517-529)
The transformer is torch's own nn.TransformerEncoder built exactly like roi_encoder.py:244-256 (batch_first=False: the
sequence axis is the class axis, length 1 at inference).
PINNING: tests/golden/lvis_roienc_*.pt hold outputs of the reference's own modules (run unmodified through
oracle/shims by oracle/make_golden.py); tests/test_oracle.py checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .meta_fcos_oracle import MetaFCOSOracle, _gn


class ROIEncoderOracle(MetaFCOSOracle):
    def __init__(self, cfg, state: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32):
        super().__init__(cfg, state, dtype)
        G = self.G
        d = int(G.TOKENIZER.FC_DIM)
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=int(G.TRANSFORMER_ENCODER.HEADS), dim_feedforward=4 * d,
                                           dropout=float(G.TRANSFORMER_ENCODER.DROPOUT))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.encoder = nn.TransformerEncoder(encoder_layer=layer, num_layers=int(G.TRANSFORMER_ENCODER.LAYERS))
        pre = "code_generator.transformer_encoder."
        self.encoder.load_state_dict({k[len(pre):]: v for k, v in self.sd.items() if k.startswith(pre)}, strict=True)
        self.encoder.to(dtype).eval()
        self.roi_bias_value = -math.log((1 - 0.01) / 0.01)   # roi_encoder.py:141-142
        self.eval_shot = int(cfg.MODEL.META_LEARN.EVAL_SHOT)

    def _re(self, name: str) -> torch.Tensor:
        return self.sd["code_generator." + name]

    def _att(self, branch: str, idx: Tuple[int, int, int, int], x: torch.Tensor) -> torch.Tensor:
        p = f"box_pooler.context_attention_module.{branch}."
        x = F.conv2d(x, self._re(p + f"{idx[0]}.weight"), self._re(p + f"{idx[0]}.bias"))
        x = F.relu(_gn(x, self._re(p + f"{idx[1]}.weight"), self._re(p + f"{idx[1]}.bias"), 32))
        x = F.conv2d(x, self._re(p + f"{idx[2]}.weight"), self._re(p + f"{idx[2]}.bias"))
        return _gn(x, self._re(p + f"{idx[3]}.weight"), self._re(p + f"{idx[3]}.bias"), 32)

    @torch.no_grad()
    def fused_roi_features(self, feats: List[torch.Tensor], boxes: torch.Tensor) -> torch.Tensor:
        """FeatureFusionModuleV2 with context_attention=True (utils.py:144-165)."""
        pooled, _ = self.roi_features(feats, boxes)
        x = F.conv2d(pooled, self._re("box_pooler.conv.0.weight"), self._re("box_pooler.conv.0.bias"), padding=1)
        x = F.relu(_gn(x, self._re("box_pooler.conv.1.weight"), self._re("box_pooler.conv.1.bias"), 32))
        context = torch.mean(torch.stack([F.adaptive_avg_pool2d(f, (7, 7)) for f in feats]), dim=0)
        local = self._att("local_att", (0, 1, 3, 4), context)
        glob = self._att("global_att", (1, 2, 4, 5), F.adaptive_avg_pool2d(context, 1))
        return x * torch.sigmoid(local + glob)

    @torch.no_grad()
    def tokens(self, x: torch.Tensor) -> torch.Tensor:
        """Tokenizer (roi_encoder.py:26-79): NUM_CONV x (conv3x3 [no bias with a norm] + GN + ReLU), flatten, FC + ReLU."""
        T = self.G.TOKENIZER
        for k in range(int(T.NUM_CONV)):
            p = f"tokenizer.conv{k + 1}."
            x = F.conv2d(x, self._re(p + "weight"), self.sd.get("code_generator." + p + "bias"), padding=1)
            if T.NORM != "":
                x = _gn(x, self._re(p + "norm.weight"), self._re(p + "norm.bias"), 32)
            x = F.relu(x)
        x = x.flatten(1)
        for k in range(int(T.NUM_FC)):
            x = F.relu(F.linear(x, self._re(f"tokenizer.fc{k + 1}.weight"), self._re(f"tokenizer.fc{k + 1}.bias")))
        return x

    def _hyper_head(self, name: str, x: torch.Tensor) -> torch.Tensor:
        n = int(self.G.HEAD.NUM_FC)
        for i in range(n):
            x = F.linear(x, self._re(f"{name}.fc{i + 1}.weight"), self._re(f"{name}.fc{i + 1}.bias"))
            if i < n - 1:
                x = F.relu(x)
        return x

    @torch.no_grad()
    def class_code(self, support_images: Sequence[torch.Tensor], boxes: torch.Tensor) -> Dict[str, torch.Tensor]:
        """run_type="meta_learn_test_support" with the ROIEncoder generator: all rows of the call are the EVAL_SHOT
        shots of bs = N / EVAL_SHOT classes (roi_encoder.py:155-163)."""
        il = self.preprocess(support_images)
        feats = self.features(il.tensor)
        n = boxes.shape[0]
        assert n % self.eval_shot == 0, f"{n} % {self.eval_shot}"
        tok = self.tokens(self.fused_roi_features(feats, boxes))
        tok = tok.view(-1, self.eval_shot, tok.shape[-1])
        tok = self.encoder(tok)                      # batch_first=False: axis 0 (classes) is the sequence
        cls_tok = tok.mean(1)
        w = self._hyper_head("weight_head", cls_tok)
        w = w.view(w.size(0), w.size(1), 1, 1)
        b = (self.roi_bias_value + self._hyper_head("bias_head", cls_tok)).view(-1)
        return {"cls_conv": w, "cls_bias": b}

    def normalize_code(self, cls_conv: torch.Tensor, cls_bias: torch.Tensor, cls_weight_norm=None):
        """No normalisation exists for this generator (the reference's normalise call raises TypeError,
        meta_one_stage_detector.py:256-259 vs roi_encoder.py:146-150): codes pass through."""
        return cls_conv, cls_bias.reshape(-1)

    @torch.no_grad()
    def head(self, feats: List[torch.Tensor], codes: Dict[str, torch.Tensor]):
        """MetaFCOSHead.forward with CondConvBlock(weight_len=256): scales[0](conv2d(x, W, b)) (head_utils.py:141-149)."""
        C = self.cfg.MODEL.FCOS
        w, b = codes["cls_conv"].to(self.dtype), codes["cls_bias"].to(self.dtype)
        scale = self._head("cond_cls_logits.scales.0.scale")
        logits, regs, ctrs, ious = [], [], [], []
        for lvl, f in enumerate(feats):
            ct = self._tower(f, "cls_tower", int(C.NUM_CLS_CONVS))
            bt = self._tower(f, "bbox_tower", int(C.NUM_BOX_CONVS))
            logits.append(F.conv2d(ct, w, b, stride=1, padding=0) * scale)
            reg = F.conv2d(bt, self._head("bbox_pred.weight"), self._head("bbox_pred.bias"), padding=1)
            if C.USE_SCALE:
                reg = reg * self._head(f"scales.{lvl}.scale")
            regs.append(F.relu(reg))
            ctrs.append(F.conv2d(bt, self._head("ctrness.weight"), self._head("ctrness.bias"), padding=1))
            ious.append(F.conv2d(bt, self._head("iou_overlap.weight"), self._head("iou_overlap.bias"), padding=1))
        return logits, regs, ctrs, ious


def build_oracle(cfg, state, dtype: torch.dtype = torch.float32) -> MetaFCOSOracle:
    """The oracle matching cfg.MODEL.META_LEARN.CODE_GENERATOR.NAME."""
    if cfg.MODEL.META_LEARN.CODE_GENERATOR.NAME == "ROIEncoder":
        return ROIEncoderOracle(cfg, state, dtype)
    return MetaFCOSOracle(cfg, state, dtype)
