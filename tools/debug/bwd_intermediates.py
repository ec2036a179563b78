#!/usr/bin/env python
"""Debugging aid: intermediates of sylph_codegen_backward (SYLPH_BWD_DEBUG_STOP) against the same algorithm in torch fp32 on
the engine's pooled ROI features.  python tools/debug/bwd_intermediates.py [case]"""
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from sylph_few_shot_detection_b200 import weights as W  # noqa: E402
from sylph_few_shot_detection_b200.modeling import build_model  # noqa: E402
from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT  # noqa: E402
from tests.cases import cfg_for, load_golden  # noqa: E402
from tests.test_gpu_training import _records  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "coco_train_2way_2shot"
g = load_golden(case)
cfg = cfg_for(g["config"], g["opts"])
state = W.synthetic_state_dict(cfg, g["seed"])
model = build_model(cfg)
model.load_state_dict(state)
eng = model.engine
support = [r for x in _records(g["items"]) for r in x["support_set"]]
shot = int(cfg.MODEL.META_LEARN.SHOT)
n = len(support); ncls = n // shot; C = 256; K = shot; L = 2; Pn = n * 49
boxes = torch.stack([r["instances"].gt_boxes.tensor[0] for r in support])
eng.extract_features(SLOT_SUPPORT, [r["image"].cuda() for r in support])
offsets = list(range(0, n + 1, shot))
raw_e = eng.generate_codes(SLOT_SUPPORT, boxes, list(range(n)), offsets)
roi = eng.export_roi_features(n).cpu()
P = "code_generator.code_generator_head."
sd = {k[len(P):]: v.float() for k, v in state.items() if k.startswith(P)}
Wt = [sd["support_set_shared_tower.0.weight"], sd["support_set_shared_tower.3.weight"]]
bt = [sd["support_set_shared_tower.0.bias"], sd["support_set_shared_tower.3.bias"]]
gw = [sd["support_set_shared_tower.1.weight"], sd["support_set_shared_tower.4.weight"]]
gb = [sd["support_set_shared_tower.1.bias"], sd["support_set_shared_tower.4.bias"]]
Wc, bc, Wb, bb = sd["support_set_cls_conv.0.weight"], sd["support_set_cls_conv.0.bias"], sd["support_set_cls_bias.0.weight"], sd["support_set_cls_bias.0.bias"]
pw, pb, cs, bs = sd["post_norm.weight"], sd["post_norm.bias"], sd["conv_scale.scale"], sd["bias_scale.scale"]
l2b = bool(cfg.MODEL.META_LEARN.CODE_GENERATOR.BIAS_L2_NORM)
gen = torch.Generator().manual_seed(7)
G = torch.randn(ncls, 257, generator=gen)
Gw, Gb = G[:, :256], G[:, 256]


def im2col(x):
    xi = x.reshape(n, 7, 7, C).permute(0, 3, 1, 2)
    return F.unfold(xi, 3, padding=1).permute(0, 2, 1).reshape(n * 49, C * 9)


def col2im(dcol):
    d = dcol.reshape(n, 49, C * 9).permute(0, 2, 1)
    return F.fold(d, (7, 7), 3, padding=1).permute(0, 2, 3, 1).reshape(n, 49, C)


def cmp(tag, got, ref):
    got, ref = got.double().cpu().reshape(-1), ref.double().reshape(-1)
    d = (got - ref).abs()
    i = int(d.argmax())
    print(f"{tag:28s} max-norm {float(d.max() / ref.abs().max()):.2e} rel-L2 {float((got - ref).norm() / ref.norm()):.2e}  worst index {i} of {ref.numel()}"
          f" got {float(got[i]):.6g} ref {float(ref[i]):.6g}")


with torch.no_grad():
    X = [roi.permute(0, 2, 3, 1).reshape(n, 49, C)]
    Y, MU, RS = [], [], []
    for i in range(L):
        y = (im2col(X[i]) @ Wt[i].reshape(C, -1).T + bt[i]).reshape(n, 49, C)
        yg = y.reshape(n, 49, 32, 8)
        mu = yg.mean(dim=(1, 3)); var = ((yg - mu[:, None, :, None]) ** 2).mean(dim=(1, 3)); rs = 1 / torch.sqrt(var + 1e-5)
        yh = ((yg - mu[:, None, :, None]) * rs[:, None, :, None]).reshape(n, 49, C)
        X.append(F.relu(yh * gw[i] + gb[i])); Y.append(y); MU.append(mu); RS.append(rs)
    colL = im2col(X[L])
    yc = colL @ Wc.reshape(C, -1).T + bc
    v = (colL @ Wb.reshape(1, -1).T + bb).reshape(n, 49)
    nrm = v.norm(dim=1, keepdim=True).clamp_min(1e-12); u = v / nrm
    shot_w = yc.reshape(n, 49, C).mean(1); shot_b = u.mean(1) if l2b else v.mean(1)
    raw_w = shot_w.reshape(ncls, K, C).mean(1); raw_b = shot_b.reshape(ncls, K).mean(1)
    raw = torch.cat([raw_w, raw_b[:, None]], dim=1)
    cmp("raw codes (engine fwd)", raw_e, raw)
    rg = raw_w.reshape(ncls, 32, 8); m8 = rg.mean(2, keepdim=True); v8 = ((rg - m8) ** 2).mean(2, keepdim=True); r8 = 1 / torch.sqrt(v8 + 1e-5)
    gh = ((rg - m8) * r8).reshape(ncls, C); gg_ = gh * pw + pb
    gn = gg_.norm(dim=1, keepdim=True).clamp_min(1e-12); l = gg_ / gn
    dl = Gw * cs
    dg = (dl - l * (l * dl).sum(1, keepdim=True)) / gn
    dgh = (dg * pw).reshape(ncls, 32, 8); ghg = gh.reshape(ncls, 32, 8)
    draw_w = (r8 * (dgh - dgh.mean(2, keepdim=True) - ghg * (dgh * ghg).mean(2, keepdim=True))).reshape(ncls, C)
    draw_b = Gb * bs
    draw = torch.cat([draw_w, draw_b[:, None]], dim=1)
    dshot_w = draw_w.repeat_interleave(K, 0) / K; dshot_b = draw_b.repeat_interleave(K, 0) / K
    dyc = (dshot_w / 49)[:, None, :].expand(n, 49, C).reshape(n * 49, C)
    du = (dshot_b / 49)[:, None].expand(n, 49)
    dv = (du - u * (u * du).sum(1, keepdim=True)) / nrm if l2b else du
    dvf = dv.reshape(n * 49, 1)
    dcolL = dyc @ Wc.reshape(C, -1) + dvf @ Wb.reshape(1, -1)
    dXL = col2im(dcolL)
    i = L - 1
    yh = ((Y[i].reshape(n, 49, 32, 8) - MU[i][:, None, :, None]) * RS[i][:, None, :, None]).reshape(n, 49, C)
    dz = dXL * ((yh * gw[i] + gb[i]) > 0)
    t = (dz * gw[i]).reshape(n, 49, 32, 8); yhg = yh.reshape(n, 49, 32, 8)
    a = t.mean(dim=(1, 3)); b2 = (t * yhg).mean(dim=(1, 3))
    dY1 = (RS[i][:, None, :, None] * (t - a[:, None, :, None] - yhg * b2[:, None, :, None])).reshape(n * 49, C)
    dbeta_part = dz.sum(dim=1)                                  # (n, C)
    dgamma_part = (dz * yh).sum(dim=1)

params = {k: state[k].cuda().float().contiguous() for k in state if k.startswith(P) and "init_norm" not in k}
read = lambda name, shape: eng.debug_read_buffer(name, shape)

os.environ["SYLPH_BWD_DEBUG_STOP"] = str(10 + L - 1)            # right after col2im of the last layer
eng.codegen_backward(offsets, raw, G, params)
torch.cuda.synchronize()
xs = read("bwd.x", (L + 1, Pn, C))
for i in range(L + 1):
    cmp(f"X[{i}]", xs[i], X[i])
ys = read("bwd.y", (L, Pn, C))
for i in range(L):
    cmp(f"Y[{i}]", ys[i], Y[i])
st = read("bwd.stats", (L, 2, n, C))
for i in range(L):
    cmp(f"mean[{i}]", st[i, 0], MU[i].repeat_interleave(8, dim=1))
    cmp(f"rstd[{i}]", st[i, 1], RS[i].repeat_interleave(8, dim=1))
cmp("draw", read("bwd.draw", (ncls, 257)), draw)
cmp("dYc", read("bwd.da", (Pn, C)), dyc)
cmp("dv", read("bwd.dv", (Pn,)), dvf)
cmp("col(X_L)", read("bwd.col", (Pn, 2304)), colL)
cmp("dcol_L", read("bwd.dcol", (Pn, 2304)), dcolL)
cmp("dX_L", read("bwd.db", (Pn, C)), dXL)
os.environ["SYLPH_BWD_DEBUG_STOP"] = str(L - 1)                 # after the GroupNorm / ReLU backward of the last layer
eng.codegen_backward(offsets, raw, G, params)
torch.cuda.synchronize()
cmp("dY_1", read("bwd.db", (Pn, C)), dY1)
parts = read("bwd.parts", (2, n, C))
cmp("dgamma partials", parts[0], dgamma_part)
cmp("dbeta partials", parts[1], dbeta_part)
