"""N > 1 host logic on CPU: world_size-2 gloo process group, class-code all-gather (the one collective of the path,
sylph/runner/meta_fcos_runner.py:381-396) and the sharding arithmetic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sylph_few_shot_detection_b200.runner import format_class_codes_shared, gather_class_code, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_classes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = []
        for c in shard_range(n_classes, world, rank):
            g = torch.Generator().manual_seed(100 + c)
            mine.append({"support_set_target": torch.tensor(c), "class_name": f"class{c}",
                         "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g),
                                        "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}})
        allc = gather_class_code(mine)
        packed = format_class_codes_shared(allc)
        q.put((rank, [int(c["support_set_target"]) for c in allc], [c["class_name"] for c in allc],
               packed["cls_conv"].clone(), packed["cls_bias"].clone()))
    finally:
        dist.destroy_process_group()


def _run(n_classes, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_classes, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out, key=lambda t: t[0])


def test_gather_class_code_world2_uneven_shards():
    n = 5  # shards of 3 and 2: the padded fixed-stride buffer must drop the padding row
    res = _run(n)
    expect_conv = torch.cat([torch.randn(1, 256, 1, 1, generator=torch.Generator().manual_seed(100 + c)) for c in range(n)])
    for rank, ids, names, conv, bias in res:
        assert ids == list(range(n)) and names == [f"class{c}" for c in range(n)]
        assert conv.shape == (n, 256, 1, 1) and bias.shape == (n,)
        assert torch.equal(conv, expect_conv)
    assert torch.equal(res[0][3], res[1][3]) and torch.equal(res[0][4], res[1][4])


def test_gather_class_code_world2_with_an_empty_rank():
    res = _run(1)  # rank 1 holds no class at all (fewer classes than ranks, as in the 5-way episode on 8 GPUs)
    for rank, ids, names, conv, bias in res:
        assert ids == [0] and conv.shape == (1, 256, 1, 1)


def test_gather_is_identity_without_a_process_group():
    codes = [{"support_set_target": 0, "class_code": {"cls_conv": torch.zeros(1, 256, 1, 1), "cls_bias": torch.zeros(1, 1, 1, 1)}}]
    assert gather_class_code(codes) is codes
