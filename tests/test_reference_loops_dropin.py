"""Drop-in check of the plugin mirror under the REFERENCE's OWN evaluation loops (build container only: needs
/root/reference).  `sylph/evaluation/meta_learn_evaluation.py` is imported unmodified -- `inference_on_support_set_dataset`
(:256-365), `inference_normalization` (:105-116), `format_class_codes_shared` (:71-103),
`inference_on_dataset_with_class_codes` (:367-470) -- and drives `sylph_few_shot_detection_b200.modeling.MetaOneStageDetector`
exactly as `MetaFCOSRunner._do_test_meta_learning` drives the reference model (meta_fcos_runner.py:451-672).

There is no GPU here, so the engine behind the mirror is a TEST DOUBLE built on the CPU oracle (test infrastructure; the
product has no CPU path).  The oracle reproduces the reference model with 0.0 deviation (tests/test_oracle.py), so any
difference from the golden outputs of the reference model would be a bug in the host mirror: run_type dispatch, record
schemas, select_a_mask, code dict shapes, code packing, detection unpacking into Instances."""
import contextlib
import sys
import types

import numpy as np
import pytest
import torch

from tests.cases import cfg_for, load_golden

pytestmark = pytest.mark.reference


class OracleBackedEngine:
    """Stands in for runtime.Engine (same method names / argument meaning), computing with the CPU oracle."""

    def __init__(self, cfg, state):
        from oracle.meta_fcos_oracle import MetaFCOSOracle
        self.orc = MetaFCOSOracle(cfg, state)
        self.device = torch.device("cpu")
        self.post_nms_topk = cfg.MODEL.FCOS.POST_NMS_TOPK_TEST
        self.slots = {}
        self.calls = []

    def extract_features(self, slot, images):
        self.calls.append(("extract_features", slot, len(images)))
        self.slots[slot] = [im.float() for im in images]

    def generate_codes(self, slot, boxes, roi_image, class_offsets, want_levels=False):
        self.calls.append(("generate_codes", slot, len(roi_image), len(class_offsets) - 1))
        rows = []
        for a, b in zip(class_offsets[:-1], class_offsets[1:]):
            code = self.orc.class_code([self.slots[slot][roi_image[i]] for i in range(a, b)], boxes[a:b])
            rows.append(torch.cat([code["cls_conv"].reshape(1, 256), code["cls_bias"].reshape(1, 1)], dim=1))
        return torch.cat(rows, dim=0)

    def normalize_codes(self, raw):
        self.calls.append(("normalize_codes", raw.shape[0]))
        rows = []
        for r in raw:
            w, b = self.orc.normalize_code(r[:256].reshape(1, 256, 1, 1), r[256:].reshape(1, 1, 1, 1))
            rows.append(torch.cat([w.reshape(1, 256), b.reshape(1, 1)], dim=1))
        return torch.cat(rows, dim=0)

    def detect_poll(self):
        pass          # the oracle never drops candidates

    def detect(self, slot, codes, out_sizes=None, max_dets=None, codes_ready=None):
        self.calls.append(("detect", slot, codes.shape[0]))
        res = self.orc.detect(self.slots[slot], {"cls_conv": codes[:, :256].reshape(-1, 256, 1, 1), "cls_bias": codes[:, 256]},
                              out_sizes)
        max_dets = max_dets or max(2 * self.post_nms_topk, 128)
        dets = torch.zeros((len(res), max_dets, 9))
        counts = torch.zeros((len(res),), dtype=torch.int32)
        for i, r in enumerate(res):
            order = torch.argsort(r["scores"], descending=True, stable=True)
            n = int(order.numel())
            dets[i, :n, 0:4] = r["boxes"][order]
            dets[i, :n, 4] = r["scores"][order]
            dets[i, :n, 5] = r["classes"][order].float()
            dets[i, :n, 6:8] = r["locations"][order]
            dets[i, :n, 8] = r["levels"][order].float()
            counts[i] = n
        return dets, counts


def _reference_evaluation_module():
    """Import sylph/evaluation/meta_learn_evaluation.py unmodified; its evaluator-side dependencies (pycocotools,
    detectron2.evaluation, d2go) are replaced by empty stand-ins -- only the inference loops are exercised."""
    from oracle import reference_loader
    reference_loader.load()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Unused:
        def __init__(self, *a, **k):
            pass

    @contextlib.contextmanager
    def inference_context(model):           # detectron2.evaluation.evaluator.inference_context
        was = model.training
        model.eval()
        yield
        model.train(was)

    class DatasetEvaluators:
        def __init__(self, evaluators):
            self.evaluators = evaluators

        def reset(self):
            pass

        def process(self, inputs, outputs):
            pass

        def evaluate(self):
            return {}

    stub("pycocotools")
    stub("pycocotools.cocoeval", COCOeval=_Unused)
    stub("pycocotools.coco", COCO=_Unused)
    stub("detectron2.evaluation")
    stub("detectron2.evaluation.coco_evaluation", COCOEvaluator=_Unused, COCOevalMaxDets=_Unused,
         _evaluate_predictions_on_coco=lambda *a, **k: None)
    stub("detectron2.evaluation.evaluator", inference_context=inference_context, DatasetEvaluators=DatasetEvaluators)
    stub("detectron2.evaluation.fast_eval_api", COCOeval_opt=_Unused)
    stub("detectron2.utils.logger", create_small_table=lambda d: str(d), log_every_n_seconds=lambda *a, **k: None)
    stub("d2go")
    stub("d2go.utils")
    stub("d2go.utils.misc", tabulate=lambda *a, **k: "")
    import os
    for pkg, path in (("sylph.data", "sylph/data"), ("sylph.data.data_injection", "sylph/data/data_injection")):
        if pkg not in sys.modules:      # their __init__ files pull in the dataset catalogs (out of scope)
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(reference_loader.REFERENCE_ROOT, path)]
            sys.modules[pkg] = m
    stub("sylph.data.data_injection.classes", COCO_BASE_CLASSES=[], COCO_NOVEL_CLASSES=[])
    import importlib
    return importlib.import_module("sylph.evaluation.meta_learn_evaluation")


class _CollectingEvaluator:
    def __init__(self):
        self.seen = []

    def reset(self):
        self.seen = []

    def process(self, inputs, outputs):
        assert len(inputs) == len(outputs)
        self.seen.extend(outputs)

    def evaluate(self):
        return {"n": len(self.seen)}


@pytest.mark.parametrize("case", ["coco_2way_2shot", "lvis_1way_3shot"])
def test_reference_evaluation_loops_drive_the_mirror_to_the_reference_results(case, tmp_path):
    from sylph_few_shot_detection_b200 import modeling as M
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    ev = _reference_evaluation_module()
    g = load_golden(case)
    cfg = cfg_for(g["config"])
    state = W.synthetic_state_dict(cfg, g["seed"])
    model = M.build_model(cfg)
    engine = OracleBackedEngine(cfg, state)
    model._state, model._engine = state, engine          # what load_state_dict does, minus the CUDA context
    for m in (model.backbone, model.proposal_generator, model.code_generator):
        m.bind_engine(engine)

    # the loaders of _do_test_meta_learning: one support set per item (batch size 1), query images batched
    support_loader = []
    for c, shots in enumerate(g["support"]):
        records = []
        for s in shots:
            h, w = s["image"].shape[-2:]
            inst = Instances((h, w))
            inst.gt_boxes = Boxes(s["box"][None])
            inst.gt_classes = torch.tensor([c])
            records.append({"image": s["image"], "instances": inst, "height": h, "width": w})
        support_loader.append([{"support_set": records, "support_set_target": torch.tensor(c), "class_name": f"class{c}"}])
    query_loader = [[{"image": q, "height": q.shape[-2], "width": q.shape[-1]} for q in g["query"]]]

    np.random.seed(0)
    out_dir = str(tmp_path / "codes")
    codes = ev.inference_on_support_set_dataset(model, support_loader, output_dir=out_dir)        # step B
    assert [c["class_name"] for c in codes] == [f"class{c}" for c in range(len(g["support"]))]
    for c, ref in zip(codes, g["raw_codes"]):
        assert c["class_code"]["cls_conv"].shape == (1, 256, 1, 1) and c["class_code"]["cls_bias"].shape == (1, 1, 1, 1)
        assert torch.equal(c["class_code"]["cls_conv"], ref["cls_conv"]) and torch.equal(c["class_code"]["cls_bias"], ref["cls_bias"])
    # the reference loop also wrote <class_name>.pth files (meta_learn_evaluation.py:316-325): our store reads them
    from sylph_few_shot_detection_b200.predictor import load_class_code_list
    stored = load_class_code_list(out_dir, [c["class_name"] for c in codes])
    assert all(torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"]) for a, b in zip(stored, codes))

    codes = ev.inference_normalization(model, codes)                                              # step D
    for c, ref in zip(codes, g["norm_codes"]):
        assert c["class_code"]["cls_bias"].shape == (1,)
        assert torch.equal(c["class_code"]["cls_conv"], ref["cls_conv"]) and torch.equal(c["class_code"]["cls_bias"], ref["cls_bias"])
    packed = ev.format_class_codes_shared(codes, torch.device("cpu"))                             # step E
    assert torch.equal(packed["cls_conv"], g["packed"]["cls_conv"]) and torch.equal(packed["cls_bias"], g["packed"]["cls_bias"])

    evaluator = _CollectingEvaluator()
    res = ev.inference_on_dataset_with_class_codes(model, query_loader, evaluator, packed)        # step F
    assert res == {"n": len(g["query"])}
    for out, ref, q in zip(evaluator.seen, g["detections"], g["query"]):
        inst = out["instances"]
        assert inst.image_size == (q.shape[-2], q.shape[-1])
        got = {(int(l), int(x), int(y), int(c)): (b, float(s)) for b, s, c, (x, y), l in
               zip(inst.pred_boxes.tensor, inst.scores, inst.pred_classes, inst.locations, inst.fpn_levels)}
        want = {(int(l), int(loc[0]), int(loc[1]), int(c)): (b, float(s)) for b, s, c, loc, l in
                zip(ref["boxes"], ref["scores"], ref["classes"], ref["locations"], ref["levels"])}
        assert set(got) == set(want) and len(inst) == len(ref["scores"])
        for k in want:
            assert torch.equal(got[k][0], want[k][0]) and got[k][1] == want[k][1]
        assert inst.pred_classes.dtype == torch.int64 and inst.fpn_levels.dtype == torch.int64
        assert bool((inst.scores[:-1] >= inst.scores[1:]).all())          # rows arrive in descending score order
    # the mirror made exactly the engine calls the C ABI offers for this loop: per class extract + generate, one
    # normalisation of all classes, one extract + detect for the query batch
    kinds = [c[0] for c in engine.calls]
    n_cls = len(g["support"])
    assert kinds == ["extract_features", "generate_codes"] * n_cls + ["normalize_codes", "extract_features", "detect"]


def test_reference_base_class_loop_equals_the_golden_accumulation():
    """`inference_on_support_set_dataset_base` (meta_learn_evaluation.py:118-254), imported unmodified, fed the seeded
    chunk codes of tests/golden/base_reduce.pt through a stand-in model: its per-rank accumulators (codes AND the
    Python-float acc_weight) equal the golden's `per_rank` entries bit for bit.  Those entries were produced by the
    RESTATED loop (oracle/base_codes_oracle.accumulate_base_codes), so this pins the restatement -- and with it the
    accumulate kernel the GPU tests compare with the same golden -- to the reference's real loop."""
    import copy
    ev = _reference_evaluation_module()
    g = load_golden("base_reduce")
    for chunks, ref in zip(g["chunks_per_rank"], g["per_rank"]):
        served = iter(chunks)

        def model(inputs, run_type=None):                 # not an nn.Module: the loop skips inference_context for it
            assert run_type == "meta_learn_test_support" and len(inputs) == 1
            return copy.deepcopy(next(served)["code"])

        loader = [[{"support_set": [], "support_set_target": torch.tensor(c["cid"]), "class_name": c["name"], "len": c["len"],
                    "total_len": c["total_len"]}] for c in chunks]
        got = ev.inference_on_support_set_dataset_base(model, loader, all_id_map={}, base_id_map={})
        assert [c["support_set_target"] for c in got] == [c["support_set_target"] for c in ref]
        for a, b in zip(got, ref):
            assert a["class_name"] == b["class_name"]
            assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
            assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
            assert a["class_code"]["acc_weight"] == b["class_code"]["acc_weight"]


# ---- MetaFCOSRunner._gather_class_code (sylph/runner/meta_fcos_runner.py:381-439).  The module around it needs d2go's
# runner stack, so the classmethod's own source is lifted out of the file (AST, nothing edited) and executed with the
# three names it uses: torch, get_world_size and the reference's reduce_class_code.
def _reference_gather_class_code():
    import ast
    import logging
    import os
    from typing import Any, Dict

    import torch.distributed as dist
    from oracle import reference_loader
    ns_ref = reference_loader.load()
    path = os.path.join(reference_loader.REFERENCE_ROOT, "sylph", "runner", "meta_fcos_runner.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "_gather_class_code")
    fn.decorator_list = []          # @classmethod -> plain function whose first argument is the class
    code = compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), path, "exec")
    env = {"torch": torch, "Dict": Dict, "Any": Any, "logger": logging.getLogger("reference"),
           "get_world_size": lambda: dist.get_world_size() if dist.is_initialized() else 1,
           "reduce_class_code": ns_ref.cg_utils.reduce_class_code}
    exec(code, env)
    return env["_gather_class_code"]


def _worker_gather(rank, world, port, q):
    import os

    import torch.distributed as dist
    from oracle import base_codes_oracle as bo
    from sylph_few_shot_detection_b200.runner import gather_class_code
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        ref_gather = _reference_gather_class_code()
        g = torch.Generator().manual_seed(700 + rank)

        def code(cid, name, w):
            return {"support_set_target": cid, "class_name": name,
                    "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, 1, 1, 1, generator=g),
                                   "acc_weight": w}}
        # uneven shards; class 7 is split over both ranks (0.4 + 0.6), class 3 lost a chunk (weights sum to 0.75)
        mine = [code(7, "seven", 0.4), code(2, "two", 1.0)] if rank == 0 else [code(3, "three", 0.75), code(7, "seven", 0.6), code(9, "nine", 1.0)]
        import copy
        want = ref_gather(None, copy.deepcopy(mine), reduce=False)
        got = gather_class_code(copy.deepcopy(mine))
        assert [(int(c["support_set_target"]), c["class_name"]) for c in got] == [(c["support_set_target"], c["class_name"]) for c in want]
        for a, b in zip(got, want):
            assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
            assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
            assert a["class_code"]["acc_weight"] == b["class_code"]["acc_weight"]          # float64 end to end
        # reduce=True: the reference merges with its own reduce_class_code; ours merges on the device (GPU tests), so here
        # the gathered list goes through the oracle's restatement, which the golden pins to the reference function
        # (reference quirk: reduce_class_code's last log line reads class_code["cls_weight_norm"], utils.py:426, which the
        # shipped configs never produce -> KeyError; a zero dummy lets the reference function run, as in oracle/make_golden.py)
        patched = copy.deepcopy(mine)
        for c in patched:
            c["class_code"]["cls_weight_norm"] = torch.zeros(1)
        with pytest.raises(KeyError):
            ref_gather(None, copy.deepcopy(mine), reduce=True)
        want_r = ref_gather(None, patched, reduce=True)
        got_r = bo.reduce_class_code(got)
        assert [int(bo._cid(c["support_set_target"])) for c in got_r] == [int(bo._cid(c["support_set_target"])) for c in want_r]
        for a, b in zip(got_r, want_r):
            assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
            assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
        q.put((rank, len(got), len(got_r)))
    finally:
        dist.destroy_process_group()


def test_gather_class_code_equals_the_reference_classmethod_on_two_ranks():
    import torch.multiprocessing as mp
    from tests.cases import fresh_rendezvous
    port = fresh_rendezvous()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gather, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, 5, 4), (1, 5, 4)]


def test_select_a_mask_equals_the_reference_function_draw_for_draw():
    """`select_a_mask` (sylph/modeling/code_generator/utils.py:27-47) next to the mirror's: same boxes for the same global
    NumPy seed over several calls (the RNG state advances identically), same all-masks mode, same ValueError on empty."""
    from oracle import reference_loader
    from oracle import upstream as up
    from sylph_few_shot_detection_b200 import modeling as M
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    ref_fn = reference_loader.load().cg_utils.select_a_mask
    g = torch.Generator().manual_seed(2)

    def make(n_boxes, kinds):
        out = []
        for cls_boxes, cls_inst in kinds:
            inst = cls_inst((64, 64))
            xy = torch.rand(n_boxes, 2, generator=torch.Generator().manual_seed(n_boxes)) * 30
            inst.gt_boxes = cls_boxes(torch.cat([xy, xy + 10], dim=1))
            out.append(inst)
        return out
    sizes = [1, 4, 2, 7, 3]
    mine_in = [make(n, [(Boxes, Instances)])[0] for n in sizes]
    ref_in = [make(n, [(up.Boxes, up.Instances)])[0] for n in sizes]
    for use_all in (False, True):
        np.random.seed(11)
        got = [M.select_a_mask(mine_in, use_all_masks=use_all) for _ in range(3)]
        np.random.seed(11)
        want = [ref_fn(ref_in, use_all_masks=use_all) for _ in range(3)]
        for a_call, b_call in zip(got, want):
            assert len(a_call) == len(b_call) == len(sizes)
            for a, b in zip(a_call, b_call):
                # the reference wraps the single box in Boxes and returns the raw tensor in all-masks mode (:41-45); the
                # mirror hands plain (k, 4) tensors to the engine either way
                assert torch.equal(a, b.tensor if hasattr(b, "tensor") else b)
    empty_mine, empty_ref = Instances((8, 8)), up.Instances((8, 8))
    empty_mine.gt_boxes, empty_ref.gt_boxes = Boxes(torch.zeros(0, 4)), up.Boxes(torch.zeros(0, 4))
    with pytest.raises(ValueError):
        ref_fn([empty_ref])
    with pytest.raises(ValueError):
        M.select_a_mask([empty_mine])
    del g


def test_format_class_codes_shared_equals_the_reference_function():
    """runner.format_class_codes_shared next to meta_learn_evaluation.format_class_codes_shared (:71-103) on code lists
    in shuffled order (the reference places class c at list index c; the mirror sorts by id: same packing)."""
    from sylph_few_shot_detection_b200.runner import format_class_codes_shared
    ev = _reference_evaluation_module()
    g = torch.Generator().manual_seed(5)
    for n, order in ((1, [0]), (5, [3, 0, 4, 1, 2]), (20, list(reversed(range(20))))):
        codes = [{"support_set_target": torch.tensor(c), "class_name": f"c{c}",
                  "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, generator=g)}}
                 for c in order]
        want = ev.format_class_codes_shared(codes, torch.device("cpu"))
        got = format_class_codes_shared(codes, device=torch.device("cpu"))
        assert set(got) == set(want) == {"cls_conv", "cls_bias"}
        assert got["cls_conv"].shape == (n, 256, 1, 1) and got["cls_bias"].shape == (n,)
        assert torch.equal(got["cls_conv"], want["cls_conv"]) and torch.equal(got["cls_bias"], want["cls_bias"])
    assert ev.format_class_codes_shared([], torch.device("cpu")) == [] and format_class_codes_shared([]) == []
    # ids that are not exactly 0..n-1 fail loudly on both sides (the reference indexes a list of n slots, then concatenates)
    def code(c):
        return {"support_set_target": torch.tensor(c), "class_name": f"c{c}",
                "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, generator=g)}}
    for bad, exc in (([code(0), code(2)], IndexError), ([code(0), code(0)], TypeError)):
        with pytest.raises(exc):
            ev.format_class_codes_shared(bad, torch.device("cpu"))
        with pytest.raises(exc):
            format_class_codes_shared(bad, device=torch.device("cpu"))


def test_runner_do_test_follows_the_reference_sequence_with_injected_loaders(tmp_path):
    """`MetaFCOSRunner.do_test(cfg, model, train_iter=None)` -> `_do_test_meta_learning` (meta_fcos_runner.py:451-672,
    674-701) with the loader / evaluator builders overridden by in-memory ones (what a deployment overrides): the
    detections the evaluator receives equal the reference model's goldens; the class codes land in
    OUTPUT_DIR/inference/default/final/<dataset>/<seed>/<class_name>.pth like the reference's."""
    import os

    from sylph_few_shot_detection_b200 import modeling as M
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.predictor import load_class_code_list
    from sylph_few_shot_detection_b200.runner import MetaFCOSRunner
    from sylph_few_shot_detection_b200.structures import Boxes, Instances
    g = load_golden("coco_2way_2shot")
    cfg = cfg_for(g["config"], ["DATASETS.TEST", ("coco_meta_val_novel",), "MODEL.META_LEARN.USE_ALL_GTS_IN_BASE_CLASSES", False,
                                "TEST.REPEAT_TEST", 2])
    cfg.OUTPUT_DIR = str(tmp_path)
    state = W.synthetic_state_dict(cfg, g["seed"])
    model = M.build_model(cfg)
    engine = OracleBackedEngine(cfg, state)
    model._state, model._engine = state, engine
    for m in (model.backbone, model.proposal_generator, model.code_generator):
        m.bind_engine(engine)

    class Runner(MetaFCOSRunner):
        def __init__(self):
            self.evaluators = []

        def build_episodic_learning_detection_test_support_set_loader(self, cfg, dataset_name, seed):
            loader = []
            for c, shots in enumerate(g["support"]):
                records = []
                for s in shots:
                    h, w = s["image"].shape[-2:]
                    inst = Instances((h, w))
                    inst.gt_boxes = Boxes(s["box"][None])
                    inst.gt_classes = torch.tensor([c])
                    records.append({"image": s["image"], "instances": inst, "height": h, "width": w})
                loader.append([{"support_set": records, "support_set_target": torch.tensor(c), "class_name": f"class{c}"}])
            return loader

        def build_episodic_learning_detection_test_query_loader(self, cfg, dataset_name):
            return [[{"image": q, "height": q.shape[-2], "width": q.shape[-1]} for q in g["query"]]]

        def get_evaluator(self, cfg, dataset_name, output_folder=None):
            self.evaluators.append(_CollectingEvaluator())
            return self.evaluators[-1]

        def thing_classes(self, cfg, dataset_name):
            return [f"class{c}" for c in range(len(g["support"]))]

    runner = Runner()
    np.random.seed(0)
    results = runner.do_test(cfg, model)
    assert list(results) == ["default", "seed0", "seed1"]                    # REPEAT_TEST repetitions at the final iteration
    assert results["seed0"] == {"coco_meta_val_novel": {"n": len(g["query"])}}
    for ev_ in runner.evaluators:
        for out, ref in zip(ev_.seen, g["detections"]):
            inst = out["instances"]
            got = {(int(l), int(x), int(y), int(c)): float(s) for s, c, (x, y), l in
                   zip(inst.scores, inst.pred_classes, inst.locations, inst.fpn_levels)}
            want = {(int(l), int(loc[0]), int(loc[1]), int(c)): float(s) for s, c, loc, l in
                    zip(ref["scores"], ref["classes"], ref["locations"], ref["levels"])}
            assert got == want
    folder = os.path.join(str(tmp_path), "inference", "default", "final", "coco_meta_val_novel", "1")
    stored = load_class_code_list(folder, ["class0", "class1"])
    for s_, ref in zip(stored, g["raw_codes"]):
        assert torch.equal(s_["class_code"]["cls_conv"], ref["cls_conv"])     # RAW codes are what is stored (:316-325)
    # the default builders say what to override instead of failing somewhere inside
    with pytest.raises(NotImplementedError, match="override"):
        MetaFCOSRunner().do_test(cfg, model)
    # in-memory episode form
    np.random.seed(0)
    res = MetaFCOSRunner().do_test(cfg, model, support_items=[b[0] for b in runner.build_episodic_learning_detection_test_support_set_loader(cfg, "", 0)],
                                   query_items=runner.build_episodic_learning_detection_test_query_loader(cfg, "")[0])
    assert [len(r["instances"]) for r in res] == [int(d["scores"].numel()) for d in g["detections"]]
