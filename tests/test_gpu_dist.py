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
