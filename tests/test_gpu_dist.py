"""Class-sharded episode over 2 GPUs with the NCCL all-gather of class codes: every rank generates
the codes of its class shard, one all_gather_into_tensor, every rank detects on its query shard with ALL codes.
Results must equal the single-GPU episode.  Skipped with fewer than 2 devices."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    from tests.cases import fresh_rendezvous      # a FileStore path: no TCP port to race for
    return fresh_rendezvous()


def _episode_inputs():
    from tests.test_gpu_cases import _images, _support_item
    ims = _images(10, 192, 256, 31)
    g = torch.Generator().manual_seed(5)
    support = []
    for c in range(3):  # 3 classes x 2 shots over 2 ranks -> shards of 2 and 1 classes
        boxes = torch.tensor([[10.0 + 20 * c, 10.0, 120.0 + 30 * c, 150.0], [30.0, 20.0 + 10 * c, 250.0, 180.0]])
        support.append(_support_item(ims[2 * c:2 * c + 2], boxes, c))
    query = [{"image": im, "height": 192, "width": 256} for im in ims[6:10]]
    return support, query


def _worker(rank, world, port, q, balance=False):
    import torch.distributed as dist
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runner import query_indices_of_rank, run_episode
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = coco_meta_fcos_cfg()
        model = build_model(cfg)
        model.pixel_mean = model.pixel_mean.to(torch.device("cuda", rank))
        model.load_state_dict(W.synthetic_state_dict(cfg, 13))
        support, query = _episode_inputs()
        res = run_episode(model, support, query, balance_queries=balance)
        mine = query_indices_of_rank(support, len(query), world, rank, balance)
        q.put((rank, mine, [(r["instances"].pred_boxes.tensor.cpu(), r["instances"].scores.cpu(),
                             r["instances"].pred_classes.cpu()) for r in res]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", [False, True])
def test_sharded_episode_matches_single_gpu(balance):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runner import run_episode
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, balance)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = coco_meta_fcos_cfg()
    model = build_model(cfg)
    model.load_state_dict(W.synthetic_state_dict(cfg, 13))
    support, query = _episode_inputs()
    ref = run_episode(model, support, query)
    seen = []
    for rank, mine, res in out:
        seen += mine
        for qi, (boxes, scores, classes) in zip(mine, res):
            r = ref[qi]["instances"]
            assert torch.equal(boxes, r.pred_boxes.tensor.cpu())
            assert torch.equal(scores, r.scores.cpu()) and torch.equal(classes, r.pred_classes.cpu())
    assert sorted(seen) == list(range(len(query)))


# ---------------------------------------------------------------------------------------------------------------
# Data-parallel meta-training step: every rank holds its own episode, the loss normalisers (positives,
# centre-ness target sum) are summed over the ranks (`reduce_sum`, fcos_outputs.py:520-523, 557-558), DDP averages the gradients.
def _train_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from tests.test_gpu_training import _records
    from tests.test_training_oracle import grad_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world, device_id=dev)
    try:
        from sylph_few_shot_detection_b200.modeling import build_model
        g, _, cfg, state = grad_case("coco_train_2way_2shot_mild")
        model = build_model(cfg)
        model.to(dev)
        model.load_state_dict(state)
        model.train()
        batched = [_records(g["items"])[rank]]                      # rank r trains on class item r
        losses = model(batched)
        sum(losses.values()).backward()
        local = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
        out = {"rank": rank, "losses": {k: float(v.detach()) for k, v in losses.items()}, "local": local}
        try:
            model.zero_grad(set_to_none=True)
            ddp = DDP(model, device_ids=[rank], find_unused_parameters=True)      # DDP_FIND_UNUSED_PARAMETERS: True in the shipped configs
            losses = ddp(batched)
            sum(losses.values()).backward()
            out["ddp"] = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
        except Exception as e:   # reported, not raised: the first phase stands on its own
            out["ddp_error"] = repr(e)
        q.put(out)
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_matches_reference_semantics():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle.make_golden import to_records
    from oracle.meta_fcos_oracle import MetaFCOSOracle
    from tests.test_training_oracle import grad_case
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=400) for _ in range(2)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the restated reference with `reduce_sum` over the two ranks: first pass for the per-rank sums, second with the totals
    g, _, cfg, state = grad_case("coco_train_2way_2shot_mild")
    orc = MetaFCOSOracle(cfg, state)
    items = to_records(g["items"])
    local_sums = []
    for r in range(2):
        _, ex = orc.training_forward([items[r]])
        local_sums.append((int(ex["num_pos"]), float(ex["ctr_targets_sum"])))
    tot_pos, tot_ctr = sum(s[0] for s in local_sums), sum(s[1] for s in local_sums)
    assert local_sums[0][0] != local_sums[1][0]                     # the normaliser differs from the single-process one

    def reduce(t):
        return torch.full_like(t, tot_pos) if t.dtype == torch.int64 else torch.full_like(t, tot_ctr)
    ref = []
    for r in range(2):
        losses, grads, _ = orc.training_grads([items[r]], world_size=2, reduce=reduce)
        ref.append(grads)
        for k, v in losses.items():
            assert abs(out[r]["losses"][k] - float(v)) <= 1e-3 * max(abs(float(v)), 1e-3), (r, k, out[r]["losses"][k], float(v))
        for k, v in grads.items():
            got = out[r]["local"][k]
            l2 = float((got - v).norm()) / max(float(v.norm()), 1e-30)
            assert l2 <= (8e-3 if "cls_tower" in k else 2e-3), (r, k, l2)
    assert "ddp_error" not in out[0] and "ddp_error" not in out[1], (out[0].get("ddp_error"), out[1].get("ddp_error"))
    for k in ref[0]:
        mean = 0.5 * (ref[0][k] + ref[1][k])
        for r in range(2):
            l2 = float((out[r]["ddp"][k] - mean).norm()) / max(float(mean.norm()), 1e-30)
            assert l2 <= (8e-3 if "cls_tower" in k else 2e-3), ("ddp", r, k, l2)
        assert torch.equal(out[0]["ddp"][k], out[1]["ddp"][k])      # one all-reduce: both ranks hold the same averaged gradient
