"""GPU parity of the episodic TRAINING forward, through the C ABI (sylph_fcos_loss_sums /
sylph_fcos_loss_finalize) and the plugin mirror (`model.train(); model(batched_inputs)`).

Tolerances: labels, target_inds and reg_targets are bit-exact against the REFERENCE golden.  The loss arithmetic is
checked twice: (a) TIGHT -- the oracle's loss restatement is fed the head outputs exported from the CUDA path, so only
the loss kernels differ (fp32 terms, fp64 accumulation): 2e-5 relative; (b) END TO END against the reference's fp32
losses: LOSS_TOL = 1e-3 relative (the north-star bar) in the default "exact" precision mode (split-fp16 operands; the
single-fp16 "fast" mode measured up to 1e-2 here: the focal loss amplifies its 1-2.5e-3 logit noise through exp())."""
import pytest
import torch

from tests.cases import cfg_for, load_golden

pytestmark = pytest.mark.gpu
LOSS_TOL = 1e-3
KERNEL_TOL = 2e-5


def _records(items):
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    def rec(r):
        h, w = r["image"].shape[-2:]
        inst = Instances((h, w))
        inst.gt_boxes = Boxes(r["boxes"].clone())
        inst.gt_classes = r["classes"].clone()
        return {"image": r["image"], "instances": inst, "height": h, "width": w}
    return [{"support_set": [rec(r) for r in it["support_set"]], "query_set": [rec(r) for r in it["query_set"]],
             "support_set_target": torch.tensor(it["support_set_target"])} for it in items]


@pytest.mark.parametrize("case", ["coco_train_2way_2shot", "lvis_train_3way_1shot_cls_only"])
def test_training_forward_matches_reference_golden(case):
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    g = load_golden(case)
    cfg = cfg_for(g["config"], g["opts"])
    state = W.synthetic_state_dict(cfg, g["seed"])
    model = build_model(cfg)
    model.load_state_dict(state)
    model.train()
    batched = _records(g["items"])
    losses, ex = model.forward_few_shot_detector_training(batched, want_targets=True)
    assert set(losses) == set(g["losses"])
    assert set(model(batched)) == set(g["losses"])                            # the public call returns the same keys
    # ---- integer / index outputs: bit-exact
    assert torch.equal(ex["labels"].cpu(), g["labels"].to(torch.int64))
    assert torch.equal(ex["target_inds"].cpu(), g["target_inds"].to(torch.int64))
    assert torch.equal(ex["reg_targets"].cpu(), g["reg_targets"])
    sums = ex["sums"].cpu()
    n_pos = int((g["labels"].to(torch.int64) != MetaFCOSOracle.BACKGROUND_ID).sum())
    assert int(sums[1]) == n_pos
    # ---- (a) loss kernels alone: oracle losses on the head outputs of the CUDA path
    orc = MetaFCOSOracle(cfg, state)
    eng = model.engine
    n_cls = len(g["items"])
    logits = [eng.export_head_output(0, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    regs = [eng.export_head_output(1, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    ctrs = [eng.export_head_output(2, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    query = [r for x in batched for r in x["query_set"]]
    targets = [int(x["support_set_target"]) for x in batched]
    ref_losses, _ = orc.fcos_losses(logits, regs, ctrs, orc.filter_gt(query, targets), targets)
    for k, v in ref_losses.items():
        got = float(losses[k])
        assert abs(got - float(v)) <= KERNEL_TOL * max(abs(float(v)), 1e-3), (k, got, float(v))
    # ---- (b) end to end against the reference's fp32 losses
    for k, v in g["losses"].items():
        got = float(losses[k])
        assert abs(got - float(v)) <= LOSS_TOL * max(abs(float(v)), 1e-3), (k, got, float(v))


def test_loss_sums_c_abi_edge_cases():
    """No ground truth at all (every location background, loc / ctr losses 0 like `reg_pred.sum() * 0`), a first box
    centred on x == 0 (centre sampling off for the image), and argument validation."""
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY, Engine
    cfg = cfg_for("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml")
    state = W.synthetic_state_dict(cfg, 4)
    eng = Engine(cfg, 0)
    eng.load_state_dict(state)
    g = torch.Generator().manual_seed(2)
    imgs = [torch.randint(0, 256, (3, 160, 224), generator=g, dtype=torch.uint8) for _ in range(2)]
    eng.extract_features(SLOT_QUERY, [im.cuda() for im in imgs])
    codes = torch.randn(3, 257, generator=g) * 0.05
    codes[:, 256] = -4.6
    orc = MetaFCOSOracle(cfg, state)
    n, _, _, lh, lw = eng.feature_shape(SLOT_QUERY)
    sizes = list(zip(lh, lw))
    cases = {
        "no_gt": [(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))] * 2,
        "cx0_quirk": [(torch.tensor([[-30.0, 10.0, 30.0, 90.0], [40.0, 30.0, 150.0, 140.0]]), torch.tensor([7, 8])),
                      (torch.tensor([[20.0, 20.0, 200.0, 150.0], [60.0, 50.0, 120.0, 110.0]]), torch.tensor([9, 7]))],
    }
    for name, gts in cases.items():
        boxes = torch.cat([b for b, _ in gts])
        classes = torch.cat([c for _, c in gts])
        offsets = [0, gts[0][0].shape[0], gts[0][0].shape[0] + gts[1][0].shape[0]]
        sums, (labels, inds, regs) = eng.fcos_loss_sums(SLOT_QUERY, codes, [7, 8, 9], boxes, classes, offsets, want_targets=True)
        lab, ind, reg, _ = orc.fcos_targets(sizes, gts)
        assert torch.equal(labels.cpu(), lab), name
        assert torch.equal(inds.cpu(), ind), name
        assert torch.equal(regs.cpu(), reg), name
        out = eng.fcos_loss_finalize(sums).cpu()
        if name == "no_gt":
            assert float(sums[1]) == 0 and float(out[1]) == 0 and float(out[2]) == 0 and float(out[0]) > 0
        else:
            first = labels[: lh[0] * lw[0]].cpu()                              # level 0, image 0
            assert bool((first == MetaFCOSOracle.BACKGROUND_ID).all())
            assert float(sums[1]) > 0
    with pytest.raises(RuntimeError, match="gt_offsets"):
        eng.fcos_loss_sums(SLOT_QUERY, codes, [7, 8, 9], torch.zeros(1, 4), torch.zeros(1, dtype=torch.int64), [0, 0, 0])


@pytest.mark.parametrize("variant", ["iou_loss", "linear_iou_loss", "no_center_sample", "radius_2p5_sizes",
                                     "focal_alpha_gamma", "focal_no_alpha"])
def test_training_forward_loss_configurations(variant):
    """Every branch of the fused target-assignment + loss kernel (IoU / linear IoU / GIoU, centre sampling on / off, another
    radius and size-of-interest table, focal alpha / gamma incl. alpha < 0) against the REFERENCE model run with the
    same config overrides (oracle/make_golden.py --train-variants-only): targets bit-exact, losses as above."""
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    base = load_golden("coco_train_2way_2shot")
    v = load_golden("coco_train_variants")["variants"][variant]
    cfg = cfg_for(base["config"], v["opts"])
    state = W.synthetic_state_dict(cfg, base["seed"])
    model = build_model(cfg)
    model.load_state_dict(state)
    model.train()
    batched = _records(base["items"])
    losses, ex = model.forward_few_shot_detector_training(batched, want_targets=True)
    assert torch.equal(ex["labels"].cpu(), v["labels"].to(torch.int64))
    assert torch.equal(ex["target_inds"].cpu(), v["target_inds"].to(torch.int64))
    assert torch.equal(ex["reg_targets"].cpu(), v["reg_targets"])
    orc = MetaFCOSOracle(cfg, state)
    eng = model.engine
    n_cls = len(base["items"])
    logits = [eng.export_head_output(0, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    regs = [eng.export_head_output(1, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    ctrs = [eng.export_head_output(2, l, SLOT_QUERY, n_cls).cpu() for l in range(5)]
    query = [r for x in batched for r in x["query_set"]]
    targets = [int(x["support_set_target"]) for x in batched]
    ref_losses, _ = orc.fcos_losses(logits, regs, ctrs, orc.filter_gt(query, targets), targets)
    assert set(losses) == set(v["losses"]) == set(ref_losses)
    for k in losses:
        got = float(losses[k])
        assert abs(got - float(ref_losses[k])) <= KERNEL_TOL * max(abs(float(ref_losses[k])), 1e-3), (k, got, float(ref_losses[k]))
        assert abs(got - float(v["losses"][k])) <= LOSS_TOL * max(abs(float(v["losses"][k])), 1e-3), (k, got, float(v["losses"][k]))


# ---------------------------------------------------------------------------------------------------------------
# Backward of the training step for the code generator (sylph_fcos_cls_loss_backward + sylph_codegen_backward behind the
# plugin's autograd hook).  GRAD_TOL: max-norm error of a gradient tensor relative to its largest entry, against the
# REFERENCE model's own `.grad` (tests/golden/train_grads.pt) and, tensor for tensor in full, against the oracle's
# autograd (pinned to that golden at 0.0 deviation by tests/test_training_oracle.py).
# The gradient of a ReLU network is a DISCONTINUOUS function of its input: an element whose pre-activation crosses 0 switches its
# whole contribution on or off.  The reference's own autograd moves by 1.6e-3 .. 2.1e-3 (max-norm per tensor) when its pooled ROI
# features move by 1e-4 relative, and by 5e-7 when they move by 1e-6 (test_reference_gradient_conditioning, CPU) -- and 1e-4 is
# what the exact-mode forward delivers (section 5 of DESIGN.md).  So end to end, against the reference model's `.grad`:
#   * tensors behind no ReLU (cls / bias convolution, post_norm, scales): GRAD_TOL = 1e-3 max-norm (measured 1e-5 .. 2e-4);
#   * the tower's tensors: relative L2 error <= GRAD_L2_TOL = 2e-3 (measured 1.4e-4 .. 1.3e-3) and max-norm <= GRAD_TOL_KINK = 1e-2
#     (measured up to 4.2e-3 on single entries of the hot case "coco_train_2way_2shot", loss_fcos_cls = 92.9).
# The kernels themselves are held to KERNEL_GRAD_TOL by the two tests below, which feed them fp32 inputs and pin the ReLU pattern.
#   * the FCOS class tower's tensors (four ReLU layers over every pyramid location; the reference's autograd moves by rel-L2 4e-3 .. 8e-3 and
#     max-norm up to 2.4e-2 under the same 1e-4 feature noise, test_reference_gradient_conditioning_class_tower): relative L2 error
#     <= TOWER_L2_TOL = 8e-3 (measured 1.6e-4 .. 2.1e-3), max-norm <= TOWER_TOL_KINK = 3e-2 (measured up to 7e-3).
GRAD_TOL = 1e-3
GRAD_L2_TOL = 2e-3
GRAD_TOL_KINK = 1e-2
TOWER_L2_TOL = 8e-3
TOWER_TOL_KINK = 3e-2
KERNEL_GRAD_TOL = 5e-5


def _train_model(case, precision=None):
    from sylph_few_shot_detection_b200.modeling import build_model
    from tests.test_training_oracle import grad_case
    g, _, cfg, state = grad_case(case)
    model = build_model(cfg)
    model.precision = precision
    model.load_state_dict(state)
    model.train()
    return g, cfg, state, model


@pytest.mark.parametrize("case", ["lvis_train_3way_1shot_cls_only", "coco_train_2way_2shot_mild", "coco_train_2way_2shot"])
def test_code_generator_gradients_match_reference(case):
    from oracle.make_golden import to_records
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from tests.test_training_oracle import check_grads_against_golden
    g, cfg, state, model = _train_model(case)
    gg = load_golden("train_grads")["cases"][case]
    names = dict(model.named_parameters())
    pre = "code_generator.code_generator_head."
    assert set(gg["grads"]) <= set(names) and all(p.is_cuda for p in names.values())
    # the code generator and (FREEZE_CLS_TOWER: False in these configurations) the FCOS class tower train; the rest is frozen
    assert all(k.startswith(("code_generator.", "proposal_generator.fcos_head.cls_tower.")) for k in names)
    assert any(k.startswith("proposal_generator.fcos_head.cls_tower.") for k in names)
    losses = model(_records(g["items"]))
    assert set(losses) == set(gg["losses"])
    sum(losses.values()).backward()
    grads = {k: p.grad for k, p in names.items() if p.grad is not None}
    assert set(grads) == set(gg["grads"])                                      # init_norm.* get none, like the reference
    tol = GRAD_TOL_KINK
    smooth = {k: v for k, v in gg["grads"].items() if "support_set_shared_tower" not in k and "cls_tower" not in k}
    check_grads_against_golden(grads, smooth, GRAD_TOL, case + " (no ReLU behind)", GRAD_TOL)
    cg = {k: v for k, v in gg["grads"].items() if "cls_tower" not in k}
    tw = {k: v for k, v in gg["grads"].items() if "cls_tower" in k}
    worst, _ = check_grads_against_golden(grads, cg, GRAD_TOL_KINK, case, GRAD_L2_TOL)
    worst_tower, _ = check_grads_against_golden(grads, tw, TOWER_TOL_KINK, case + " (class tower)", TOWER_L2_TOL)
    worst = max(worst, worst_tower)
    # gradient with respect to the final class codes
    gc = model._last_grad_codes.cpu()
    ref_w = gg["grad_codes"]["cls_conv"].reshape(-1, 256)
    ref_b = gg["grad_codes"]["cls_bias"].reshape(-1)
    assert float((gc[:, :256] - ref_w).abs().max()) <= GRAD_TOL * float(ref_w.abs().max())      # no ReLU between the loss and the codes
    assert float((gc[:, 256] - ref_b).abs().max()) <= GRAD_TOL * float(ref_b.abs().max())
    # every tensor in full against the oracle's autograd
    orc = MetaFCOSOracle(cfg, state)
    _, ograds, _ = orc.training_grads(to_records(g["items"]))
    for k, v in ograds.items():
        err = float((grads[k].cpu() - v).abs().max()) / max(float(v.abs().max()), 1e-12)
        l2 = float((grads[k].cpu() - v).norm()) / max(float(v.norm()), 1e-30)
        if "cls_tower" in k:
            assert err <= TOWER_TOL_KINK and l2 <= TOWER_L2_TOL, (k, err, l2)
        else:
            assert err <= tol and l2 <= GRAD_L2_TOL, (k, err, l2)
        worst = max(worst, err)
    print(f"[{case}] worst relative gradient error {worst:.2e}")


@pytest.mark.parametrize("case,opts", [
    ("lvis_train_3way_1shot_cls_only", []),
    ("coco_train_2way_2shot", []),
    ("coco_train_2way_2shot", ["MODEL.META_LEARN.CODE_GENERATOR.BIAS_L2_NORM", True]),
    ("coco_train_2way_2shot", ["MODEL.META_LEARN.SHOT", 1]),
    ("coco_train_2way_2shot", ["MODEL.META_LEARN.SHOT", 4, "MODEL.META_LEARN.CODE_GENERATOR.POST_NORM", ""]),
    ("synthetic_15_rois", ["MODEL.META_LEARN.SHOT", 5]),
    ("synthetic_15_rois", ["MODEL.META_LEARN.SHOT", 3, "MODEL.META_LEARN.CODE_GENERATOR.TOWER_LAYERS", [["GN", "ReLU"], ["GN", "ReLU"], ["GN", "ReLU"]]]),
])
def test_codegen_backward_kernels_alone(case, opts):
    """sylph_codegen_backward against autograd through the oracle's code generator on the SAME fp32 inputs: the pooled ROI
    features exported from the engine, the oracle's raw codes and a random upstream gradient -- only the backward kernels
    (fp32 re-evaluation of the tower, GroupNorm / ReLU / pool / L2 / normalisation backward, the three GEMM forms) differ.
    Configuration variants: BIAS_L2_NORM on / off, 1 / 2 / 4 shots per class, POST_NORM on / off."""
    import torch.nn.functional as F
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    if case == "synthetic_15_rois":      # the ROI count of a meta-training batch (3 classes x 5 shots): 735 pixel rows, 12 GEMM row tiles
        cfg = cfg_for("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", list(opts))
        state = W.synthetic_state_dict(cfg, 3)
        gi = torch.Generator().manual_seed(15)
        images = [torch.randint(0, 256, (3, 160, 224), generator=gi, dtype=torch.uint8) for _ in range(15)]
        x0 = torch.rand(15, generator=gi) * 100
        y0 = torch.rand(15, generator=gi) * 60
        boxes = torch.stack([x0, y0, x0 + 20 + torch.rand(15, generator=gi) * 100, y0 + 20 + torch.rand(15, generator=gi) * 70], dim=1)
    else:
        g = load_golden(case)
        cfg = cfg_for(g["config"], list(g["opts"]) + list(opts))
        state = W.synthetic_state_dict(cfg, g["seed"])
        support = [r for x in _records(g["items"]) for r in x["support_set"]]
        images = [r["image"] for r in support]
        boxes = torch.stack([r["instances"].gt_boxes.tensor[0] for r in support])
    model = build_model(cfg)
    model.load_state_dict(state)
    eng = model.engine
    shot = int(cfg.MODEL.META_LEARN.SHOT)
    n, n_cls = len(images), len(images) // shot
    eng.extract_features(SLOT_SUPPORT, [im.cuda() for im in images])
    offsets = list(range(0, n + 1, shot))
    eng.generate_codes(SLOT_SUPPORT, boxes, list(range(n)), offsets)
    roi = eng.export_roi_features(n).cpu()
    orc = MetaFCOSOracle(cfg, state)
    keys = orc.trainable_code_generator_keys()
    params = {k: state[k].cuda().float().contiguous() for k in keys}
    gen = torch.Generator().manual_seed(7)
    G = torch.randn(n_cls, 257, generator=gen)
    L = len(cfg.MODEL.META_LEARN.CODE_GENERATOR.TOWER_LAYERS)

    def oracle_grads(masks):
        """autograd through per_shot_codes -> class mean -> process_codes_training; `masks` (one (n, 256, 7, 7) bool tensor per tower
        layer) pins every ReLU to x * mask, None = the plain ReLU."""
        leaves = {k: orc.sd[k].detach().clone().requires_grad_(True) for k in keys}
        saved = {k: orc.sd[k] for k in keys}
        orc.sd.update(leaves)
        pre = {}
        if masks is not None:
            def tower(x):
                for i in range(L):
                    x = F.conv2d(x, orc._cg(f"support_set_shared_tower.{3 * i}.weight"), orc._cg(f"support_set_shared_tower.{3 * i}.bias"), padding=1)
                    x = F.group_norm(x, 32, orc._cg(f"support_set_shared_tower.{3 * i + 1}.weight"), orc._cg(f"support_set_shared_tower.{3 * i + 1}.bias"), 1e-5)
                    pre[i] = x.detach()
                    x = x * masks[i]
                return x
            orc._cg_tower = tower
        try:
            with torch.enable_grad():
                w, b = MetaFCOSOracle.per_shot_codes.__wrapped__(orc, roi)
                raw_w = w.view(n_cls, shot, *w.shape[1:]).mean(dim=1)
                raw_b = b.view(n_cls, shot, 1, 1, 1).mean(dim=1)
                fw, fb = orc.process_codes_training(raw_w, raw_b)
                ((fw.reshape(n_cls, 256) * G[:, :256]).sum() + (fb.reshape(-1) * G[:, 256]).sum()).backward()
            grads = {k: leaves[k].grad.detach() for k in keys if leaves[k].grad is not None}
        finally:
            orc.sd.update(saved)
            if masks is not None:
                del orc._cg_tower
        raw = torch.cat([raw_w.detach().reshape(n_cls, 256), raw_b.detach().reshape(n_cls, 1)], dim=1)
        return grads, raw, pre

    _, raw, _ = oracle_grads(None)
    got = eng.codegen_backward(offsets, raw, G, params)
    # A ReLU is not differentiable at 0: two fp32 evaluations of the same tower that differ in the last bits may put an element
    # on different sides, and ONE such element moves single gradient entries by ~4e-3 of the largest (measured; the reference's
    # own autograd moves by 2e-3 when its input moves by 1e-4, tests/test_training_oracle.py).  So the reference is
    # evaluated with the ReLU pattern of the engine's fp32 re-evaluation (its layer inputs X_1 .. X_L, read back), and the
    # patterns themselves may only differ where the pre-activation is within rounding distance of 0.
    xs = eng.debug_read_buffer("bwd.x", (L + 1, n, 7, 7, 256)).cpu()
    masks = [(xs[i + 1] > 0).permute(0, 3, 1, 2) for i in range(L)]
    ref, _, pre = oracle_grads(masks)
    flips = 0
    for i in range(L):
        differ = masks[i] != (pre[i] > 0)
        flips += int(differ.sum())
        assert float(pre[i][differ].abs().max()) < 1e-4 if differ.any() else True, "ReLU patterns differ away from 0"
    assert flips <= 8, flips
    errs = {k.split("head.")[1]: (float((got[k].cpu() - ref[k]).abs().max()) / max(float(ref[k].abs().max()), 1e-12),
                                  float((got[k].cpu() - ref[k]).norm()) / max(float(ref[k].norm()), 1e-30)) for k in ref}
    print(f"[{case} {opts}] codegen backward kernels vs fp32 autograd ({flips} ReLU ties), max-norm / rel-L2: " +
          ", ".join(f"{k} {a:.1e}/{b:.1e}" for k, (a, b) in errs.items()))
    bad = {k: v for k, v in errs.items() if v[0] > KERNEL_GRAD_TOL}
    assert not bad, bad


def test_cls_loss_backward_kernel_alone():
    """sylph_fcos_cls_loss_backward against the focal-loss gradient autograd gives on the logits EXPORTED from the engine.
    Bias gradient: sum over locations of d loss / d logit.  Convolution gradient by the adjoint identity
    <d loss / d cls_conv[c], delta> = sum over locations of d loss / d logit[loc, c] * <tower[loc], delta>, where <tower[loc], delta> is
    what the engine's own conditional convolution returns for the code `delta` (bias 0) -- so the class-tower output
    never has to leave the device."""
    from oracle import upstream as up
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    g, cfg, state, model = _train_model("coco_train_2way_2shot")
    eng = model.engine
    batched = _records(g["items"])
    query = [r for x in batched for r in x["query_set"]]
    targets = [int(x["support_set_target"]) for x in batched]
    gts = model._get_gt(query, support_set_targets=[x["support_set_target"] for x in batched])
    eng.extract_features(SLOT_QUERY, [r["image"].cuda() for r in query])
    gen = torch.Generator().manual_seed(3)
    n_cls = len(targets)
    codes = torch.randn(n_cls, 257, generator=gen) * 0.05
    codes[:, 256] = torch.tensor([-2.0, -4.0])[:n_cls]
    gb = torch.cat([x.gt_boxes.tensor.reshape(-1, 4) for x in gts])
    gc = torch.cat([x.gt_classes.reshape(-1) for x in gts])
    off = [0]
    for x in gts:
        off.append(off[-1] + len(x.gt_classes))

    def flat(which, c):
        return torch.cat([eng.export_head_output(which, l, SLOT_QUERY, n_cls).permute(0, 2, 3, 1).reshape(-1, c) for l in range(5)])

    sums, (labels, _, _) = eng.fcos_loss_sums(SLOT_QUERY, codes, targets, gb, gc, off, want_targets=True)
    up_grad = torch.tensor([1.7], device="cuda")
    got = eng.fcos_cls_loss_backward(SLOT_QUERY, n_cls, targets, labels, sums, grad_loss=up_grad).double().cpu()
    logits = flat(0, n_cls).double().cpu().requires_grad_(True)
    st = torch.tensor(targets).view(1, -1)
    tgt = (st == labels.cpu()[:, None]).double()
    n_pos = max(float((labels != 100000).sum()), 1.0)
    C = cfg.MODEL.FCOS
    loss = up.sigmoid_focal_loss(logits, tgt, alpha=C.LOSS_ALPHA, gamma=C.LOSS_GAMMA, reduction="sum") / n_pos
    (1.7 * loss).backward()
    G = logits.grad                                                           # (locations, classes)
    ref_b = G.sum(dim=0)
    assert float((got[:, 256] - ref_b).abs().max()) <= KERNEL_GRAD_TOL * float(ref_b.abs().max())
    # adjoint identity with n_cls random directions
    delta = torch.randn(n_cls, 257, generator=gen)
    delta[:, 256] = 0.0
    eng.fcos_loss_sums(SLOT_QUERY, delta, targets, gb, gc, off)
    proj = flat(0, n_cls).double().cpu()                                      # <tower[loc], delta_j>
    ref = G.t() @ proj                                                        # (class c, direction j)
    mine = got[:, :256] @ delta[:, :256].double().t()
    assert float((mine - ref).abs().max()) <= 2e-4 * float(ref.abs().max()), (mine, ref)


def test_code_generator_backward_is_deterministic_and_guards_stale_buffers():
    g, cfg, state, model = _train_model("lvis_train_3way_1shot_cls_only")
    batched = _records(g["items"])
    runs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        losses = model(batched)
        sum(losses.values()).backward()
        runs.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert runs[0].keys() == runs[1].keys() and len(runs[0]) == 32      # 16 code-generator + 16 class-tower tensors
    for k in runs[0]:
        assert torch.equal(runs[0][k], runs[1][k]), k                          # fixed-order reductions, no atomics
    # an upstream factor scales every gradient (the hook receives d total / d loss_fcos_cls on the device)
    model.zero_grad(set_to_none=True)
    (2.0 * model(batched)["loss_fcos_cls"]).backward()
    k = "code_generator.code_generator_head.support_set_cls_conv.0.weight"
    got = dict(model.named_parameters())[k].grad
    assert float((got - 2.0 * runs[0][k]).abs().max()) <= 1e-5 * float(runs[0][k].abs().max())
    # backward of an episode whose engine buffers a later forward has overwritten must fail loudly
    first = model(batched)
    model(batched)
    with pytest.raises(RuntimeError, match="before the next forward"):
        first["loss_fcos_cls"].backward()
    # no_grad forward: plain loss tensors, no hook
    with torch.no_grad():
        assert not model(batched)["loss_fcos_cls"].requires_grad
    # reloading a checkpoint into the training model (a fresh engine underneath) keeps the parameters and the training mode
    ids = [id(p) for p in model.parameters()]
    model.load_state_dict(state)
    assert [id(p) for p in model.parameters()] == ids
    model.zero_grad(set_to_none=True)
    sum(model(batched).values()).backward()
    for k, p in model.named_parameters():
        if k in runs[0]:
            assert torch.equal(p.grad, runs[0][k]), k


def test_optimizer_step_reaches_the_engine():
    """One SGD step on the plugin's parameters: the next forward runs with the stepped weights (sylph_update_code_generator)
    and agrees with the oracle evaluated on the stepped state_dict; state_dict() returns the live values."""
    from oracle.make_golden import to_records
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    g, cfg, state, model = _train_model("lvis_train_3way_1shot_cls_only")
    batched = _records(g["items"])
    opt = torch.optim.SGD(model.parameters(), lr=1e-4)   # the oracle's loss on this episode: 1.692 -> 1.583 at this step size
    l0 = model(batched)["loss_fcos_cls"]
    l0.backward()
    opt.step()
    l1 = float(model(batched)["loss_fcos_cls"])
    assert l1 != float(l0)
    stepped = model.state_dict()
    k = "code_generator.code_generator_head.support_set_cls_conv.0.weight"
    assert not torch.equal(stepped[k], state[k])
    trained = ("code_generator.", "proposal_generator.fcos_head.cls_tower.")
    assert all(torch.equal(stepped[q], state[q]) for q in state if not q.startswith(trained))       # the rest of the detector is frozen
    assert not torch.equal(stepped["proposal_generator.fcos_head.cls_tower.0.weight"], state["proposal_generator.fcos_head.cls_tower.0.weight"])
    orc = MetaFCOSOracle(cfg, {q: v.clone() for q, v in stepped.items()})
    ref, _ = orc.training_forward(to_records(g["items"]))
    assert abs(l1 - float(ref["loss_fcos_cls"])) <= LOSS_TOL * abs(float(ref["loss_fcos_cls"])), (l1, float(ref["loss_fcos_cls"]))
    # the plain SGD step lowers this episode's loss (sanity of the gradient's sign)
    assert l1 < float(l0), (l1, float(l0))
    # a second step, then straight to eval: the inference entry points see the stepped weights too
    model.zero_grad(set_to_none=True)
    model(batched)["loss_fcos_cls"].backward()
    opt.step()
    model.eval()
    item = {"support_set": batched[0]["support_set"], "support_set_target": batched[0]["support_set_target"]}
    import numpy as np
    np.random.seed(0)
    code = model([item], run_type="meta_learn_test_support")
    orc2 = MetaFCOSOracle(cfg, {q: v.clone() for q, v in model.state_dict().items()})
    ref_code = orc2.class_code([r["image"] for r in item["support_set"]],
                               torch.stack([r["instances"].gt_boxes.tensor[0] for r in item["support_set"]]))
    got_w = code["cls_conv"].cpu().reshape(-1)
    assert float((got_w - ref_code["cls_conv"].reshape(-1)).abs().max()) <= 1e-3 * float(ref_code["cls_conv"].abs().max())
    orc1 = MetaFCOSOracle(cfg, {q: v.clone() for q, v in stepped.items()})
    old_code = orc1.class_code([r["image"] for r in item["support_set"]],
                               torch.stack([r["instances"].gt_boxes.tensor[0] for r in item["support_set"]]))
    assert float((got_w - old_code["cls_conv"].reshape(-1)).abs().max()) > 1e-3 * float(ref_code["cls_conv"].abs().max())   # not the weights of one step ago


def test_backward_c_abi_validation():
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    g, cfg, state, model = _train_model("lvis_train_3way_1shot_cls_only")
    eng = model.engine
    labels = torch.zeros(10, dtype=torch.int64, device="cuda")
    sums = torch.zeros(5, dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="holds no features"):
        eng.fcos_cls_loss_backward(SLOT_QUERY, 3, [1, 2, 3], labels, sums)
    eng.extract_features(SLOT_QUERY, [torch.zeros(3, 64, 64, dtype=torch.uint8, device="cuda")])
    with pytest.raises(RuntimeError, match="last head pass"):
        eng.fcos_cls_loss_backward(SLOT_QUERY, 3, [1, 2, 3], labels, sums)
    with pytest.raises(RuntimeError, match="last sylph_generate_codes"):
        eng.codegen_backward([0, 1, 2], torch.zeros(2, 257), torch.zeros(2, 257), {})
    # the class tower's backward needs the activations of a head pass in training mode
    live = {k: p.detach() for k, p in model.named_parameters()}
    eng.set_training(False)
    batched = _records(g["items"])
    with torch.no_grad():
        model(batched)
    with pytest.raises(RuntimeError, match="training mode"):
        eng.cls_tower_backward(SLOT_QUERY, torch.zeros(3, 257), [40, 2, 17], labels, sums, live)
    eng.set_training(True)
    with pytest.raises(RuntimeError, match="missing parameter"):
        with torch.no_grad():
            losses, ex = model.forward_few_shot_detector_training(batched, want_targets=True)
        eng.cls_tower_backward(SLOT_QUERY, torch.zeros(3, 257), [40, 2, 17], ex["labels"], ex["sums"], {})


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_device_weight_refresh_is_bit_identical_to_host_preparation(precision):
    """sylph_update_code_generator_device (kernels pack the optimiser's device tensors into the operand layouts) against
    sylph_update_code_generator (the host preparation sylph_finalize_weights uses): the raw class codes of the same support
    set are bit-identical, and differ from the codes of the previous weights."""
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.runtime import SLOT_SUPPORT
    g = load_golden("coco_train_2way_2shot")
    cfg = cfg_for(g["config"], g["opts"])
    state = W.synthetic_state_dict(cfg, g["seed"])
    model = build_model(cfg)
    model.precision = precision
    model.load_state_dict(state)
    eng = model.engine
    support = [r for x in _records(g["items"]) for r in x["support_set"]]
    boxes = torch.stack([r["instances"].gt_boxes.tensor[0] for r in support])
    eng.extract_features(SLOT_SUPPORT, [r["image"].cuda() for r in support])
    n = len(support)

    def codes():
        raw = eng.generate_codes(SLOT_SUPPORT, boxes, list(range(n)), [0, n // 2, n])
        return torch.cat([raw, eng.normalize_codes(raw)], dim=1).cpu()

    c0 = codes()
    gen = torch.Generator().manual_seed(5)
    stepped = {k: (v + 0.01 * v.abs().mean() * torch.randn(v.shape, generator=gen)).float() for k, v in state.items()
               if k.startswith("code_generator.")}
    eng.update_code_generator_device({k: v.cuda().contiguous() for k, v in stepped.items()})
    c_dev = codes()
    assert not torch.equal(c_dev, c0)
    eng.update_code_generator(state)          # back to the loaded weights through the host path
    assert torch.equal(codes(), c0)
    eng.update_code_generator(stepped)
    assert torch.equal(codes(), c_dev)


@pytest.mark.parametrize("precision,tol", [("exact", 5e-4), ("fast", 1e-2)])
def test_cls_tower_backward_kernels_alone(precision, tol):
    """sylph_cls_tower_backward against fp32 autograd on the engine's OWN inputs and ReLU pattern: the pyramid exported from the
    engine feeds a torch class tower whose ReLUs are pinned to the pattern of the engine's saved activations, the engine's final
    codes condition the classifier, the focal loss is differentiated by autograd.  What is left are the kernels: tower-output
    gradient, GroupNorm / ReLU backward over the planes, the scaled fp16 (hi | lo) dY, the MN-major tcgen05 weight gradient, the
    input-gradient convolutions (three products per multiply).  "fast": single fp16 operands (11-bit dY, one product)."""
    import torch.nn.functional as F
    from oracle import upstream as up
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from sylph_few_shot_detection_b200.runtime import SLOT_QUERY
    g, cfg, state, model = _train_model("lvis_train_3way_1shot_cls_only", precision)
    eng = model.engine
    assert eng.precision == precision
    batched = _records(g["items"])
    losses = model(batched)
    sum(losses.values()).backward()
    pre = "proposal_generator.fcos_head.cls_tower."
    got = {k: p.grad.cpu() for k, p in model.named_parameters() if k.startswith(pre)}
    assert len(got) == 16
    n, _, _, lh, lw = eng.feature_shape(SLOT_QUERY)
    feats = [eng.export_features(SLOT_QUERY, l).cpu() for l in range(5)]
    codes = model._last_final_codes.cpu()
    n_cls = codes.shape[0]
    # ReLU pattern of the engine's saved layer outputs (planes "det.cls_x<i>": rows [256 hi | 256 lo], level-major, 1-pixel border)
    r128 = lambda v: (v + 127) // 128 * 128
    rows_per = [r128((h + 2) * (w + 2)) for h, w in zip(lh, lw)]
    level_row0 = [0]
    for rp in rows_per:
        level_row0.append(level_row0[-1] + n * rp)
    L = int(cfg.MODEL.FCOS.NUM_CLS_CONVS)
    masks = []
    for i in range(L):
        ld = 512 if precision == "exact" else 256
        planes = eng.debug_read_buffer(f"det.cls_x{i}", (level_row0[5], ld), torch.float16).cpu().float()
        val = planes[:, :256] + planes[:, 256:] if precision == "exact" else planes
        per_level = []
        for l, (h, w) in enumerate(zip(lh, lw)):
            blk = val[level_row0[l]:level_row0[l + 1]].reshape(n, rows_per[l], 256)[:, :(h + 2) * (w + 2)].reshape(n, h + 2, w + 2, 256)
            per_level.append((blk[:, 1:h + 1, 1:w + 1] > 0).permute(0, 3, 1, 2))
        masks.append(per_level)
    orc = MetaFCOSOracle(cfg, state)
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in got}
    query = [r for x in batched for r in x["query_set"]]
    targets = [int(x["support_set_target"]) for x in batched]
    gts = orc.filter_gt(query, targets)
    labels, _, _, _ = orc.fcos_targets([(h, w) for h, w in zip(lh, lw)], gts)
    flips = 0
    logits = []
    with torch.enable_grad():
        for l in range(5):
            x = feats[l]
            for i in range(L):
                x = F.conv2d(x, leaves[f"{pre}{3 * i}.weight"], leaves[f"{pre}{3 * i}.bias"], padding=1)
                x = F.group_norm(x, 32, leaves[f"{pre}{3 * i + 1}.weight"], leaves[f"{pre}{3 * i + 1}.bias"], 1e-5)
                flips += int(((x.detach() > 0) != masks[i][l]).sum())
                x = x * masks[i][l]
            logits.append(F.conv2d(x, codes[:, :256].reshape(n_cls, 256, 1, 1), codes[:, 256]))
        pred = torch.cat([t.permute(0, 2, 3, 1).reshape(-1, n_cls) for t in logits])
        tgt = (torch.tensor(targets).view(1, -1) == labels[:, None]).float()
        n_pos = max(float((labels != MetaFCOSOracle.BACKGROUND_ID).sum()), 1.0)
        C = cfg.MODEL.FCOS
        (up.sigmoid_focal_loss(pred, tgt, alpha=C.LOSS_ALPHA, gamma=C.LOSS_GAMMA, reduction="sum") / n_pos).backward()
    total = sum(m.numel() for per in masks for m in per)
    assert flips <= (2e-3 if precision == "exact" else 2e-2) * total, (flips, total)   # the patterns differ only near 0 (forward noise)
    errs = {k[len(pre):]: (float((got[k] - leaves[k].grad).abs().max() / leaves[k].grad.abs().max()),
                           float((got[k] - leaves[k].grad).norm() / leaves[k].grad.norm())) for k in got}
    print(f"[class tower kernels alone, {precision}, {flips} of {total} ReLU decisions differ from torch's own forward] max-norm / rel-L2: " +
          ", ".join(f"{k} {a:.1e}/{b:.1e}" for k, (a, b) in errs.items()))
    bad = {k: v for k, v in errs.items() if v[0] > tol or v[1] > tol}
    assert not bad, bad
