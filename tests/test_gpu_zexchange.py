"""Class-code exchange over NVLink peer memory (`sylph_normalize_codes_exchange`): the normalisation kernel
stores every finished row into the exchange buffer of every rank and a small kernel on each rank waits for the rows of
the episode.  One GPU: the kernel pair against `sylph_normalize_codes` (bit-identical rows, both buffer halves, state
advance).  Two GPUs: the sharded episode with `exchange="peer"` against the single-GPU episode (skipped on a 1-GPU box)."""
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_gpu_cases import _setup

pytestmark = pytest.mark.gpu


def test_exchange_kernels_match_normalize_codes_one_rank():
    cfg, state, model, orc = _setup()
    eng = model.engine
    eng.exchange_setup(None, max_classes=64)
    try:
        g = torch.Generator().manual_seed(3)
        total = 0
        for n in (5, 1, 20, 64, 7):            # odd count of calls: both halves of the double buffer, several sizes
            raw = torch.randn((n, 257), generator=g).cuda()
            want = eng.normalize_codes(raw)
            got = eng.normalize_codes_exchange(raw, 0, n)
            assert got.shape == (n, 257)
            assert torch.equal(got, want)
            total += n
        timed_out, rows = eng.exchange_status()
        assert not timed_out and rows == total
        eng.exchange_poll()                    # nothing to report
        with pytest.raises(RuntimeError):
            eng.normalize_codes_exchange(torch.zeros((65, 257)).cuda(), 0, 65)     # more classes than the buffer holds
        with pytest.raises(RuntimeError):
            eng.normalize_codes_exchange(torch.zeros((4, 257)).cuda(), 2, 5)       # shard sticks out of the class range
    finally:
        eng.exchange_teardown(None)


def test_exchange_times_out_instead_of_hanging(monkeypatch):
    """A rank that waits for rows nobody sends gives up after SYLPH_EXCHANGE_TIMEOUT_MS and raises the status flag."""
    monkeypatch.setenv("SYLPH_EXCHANGE_TIMEOUT_MS", "50")
    cfg, state, model, orc = _setup()
    eng = model.engine
    eng.exchange_setup(None, max_classes=8)
    try:
        eng.normalize_codes_exchange(torch.zeros((2, 257)).cuda(), 0, 3)   # 3 classes announced, 2 delivered
        timed_out, rows = eng.exchange_status()
        assert timed_out and rows == 2
        with pytest.raises(RuntimeError, match="gave up"):
            eng.exchange_poll()                # the flag has followed the rows to pinned host memory
        with pytest.raises(RuntimeError, match="gave up"):
            eng.normalize_codes_exchange(torch.zeros((1, 257)).cuda(), 0, 1)   # and the next exchange refuses to run
    finally:
        eng.exchange_teardown(None)


def _worker(rank, world, port, q):
    import os

    import torch.distributed as dist
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runner import query_indices_of_rank, run_episode
    from tests.test_gpu_dist import _episode_inputs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = coco_meta_fcos_cfg()
        model = build_model(cfg)
        model.pixel_mean = model.pixel_mean.to(torch.device("cuda", rank))
        model.load_state_dict(W.synthetic_state_dict(cfg, 13))
        support, query = _episode_inputs()
        out = []
        for _ in range(3):                      # three episodes: both buffer halves and the episode counters
            res = run_episode(model, support, query, exchange="peer")
            out.append([(r["instances"].pred_boxes.tensor.cpu(), r["instances"].scores.cpu(), r["instances"].pred_classes.cpu())
                        for r in res])
        timed_out, rows = model.engine.exchange_status()
        mine = query_indices_of_rank(support, len(query), world, rank, False)
        q.put((rank, mine, out, timed_out, rows))
        model.engine.exchange_teardown(None)
    finally:
        dist.destroy_process_group()


def test_sharded_episode_with_peer_exchange_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sylph_few_shot_detection_b200 import weights as W
    from sylph_few_shot_detection_b200.modeling import build_model
    from sylph_few_shot_detection_b200.presets import coco_meta_fcos_cfg
    from sylph_few_shot_detection_b200.runner import run_episode
    from tests.test_gpu_dist import _episode_inputs, _free_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = coco_meta_fcos_cfg()
    model = build_model(cfg)
    model.load_state_dict(W.synthetic_state_dict(cfg, 13))
    support, query = _episode_inputs()
    ref = run_episode(model, support, query)
    for rank, mine, episodes, timed_out, rows in got:
        assert not timed_out and rows == 3 * len(support)
        for res in episodes:
            for qi, (boxes, scores, classes) in zip(mine, res):
                r = ref[qi]["instances"]
                assert torch.equal(boxes, r.pred_boxes.tensor.cpu())
                assert torch.equal(scores, r.scores.cpu()) and torch.equal(classes, r.pred_classes.cpu())
