"""N > 1 host logic on CPU: world_size-2 gloo process group, class-code all-gather (the one collective of the path,
sylph/runner/meta_fcos_runner.py:381-396) and the sharding arithmetic."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sylph_few_shot_detection_b200.runner import (format_class_codes_shared, gather_class_code,
                                                   gather_class_code_known_shards, shard_range)


def _free_port():
    from tests.cases import fresh_rendezvous
    return fresh_rendezvous()


def _worker(rank, world, port, n_classes, q):
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        mine = []
        for c in shard_range(n_classes, world, rank):
            g = torch.Generator().manual_seed(100 + c)
            mine.append({"support_set_target": torch.tensor(c), "class_name": f"class{c}",
                         "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g),
                                        "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}})
        allc = gather_class_code(mine)
        packed = format_class_codes_shared(allc)
        # the one-collective fast path of run_episode (shard sizes / ids / names known on every rank) must agree
        counts = [len(shard_range(n_classes, world, r)) for r in range(world)]
        meta = [(torch.tensor(c), f"class{c}") for c in range(n_classes)]
        fast = gather_class_code_known_shards(mine, counts, meta)
        assert [int(c["support_set_target"]) for c in fast] == [int(c["support_set_target"]) for c in allc]
        assert [c["class_name"] for c in fast] == [c["class_name"] for c in allc]
        for a, b in zip(fast, allc):
            assert torch.equal(a["class_code"]["cls_conv"], b["class_code"]["cls_conv"])
            assert torch.equal(a["class_code"]["cls_bias"], b["class_code"]["cls_bias"])
        q.put((rank, [int(c["support_set_target"]) for c in allc], [c["class_name"] for c in allc],
               packed["cls_conv"].clone(), packed["cls_bias"].clone()))
    finally:
        dist.destroy_process_group()


def _worker_base(rank, world, port, q):
    """Base-class path: both ranks hold a PARTIAL sum of class 7 (acc_weight 0.4 / 0.6); rank 1 also holds class 3."""
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        g = torch.Generator().manual_seed(500 + rank)
        mine = [{"support_set_target": 7, "class_name": "seven",
                 "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, 1, 1, 1, generator=g),
                                "acc_weight": 0.4 if rank == 0 else 0.6}}]
        if rank == 1:
            mine.append({"support_set_target": 3, "class_name": "three",
                         "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g),
                                        "cls_bias": torch.randn(1, 1, 1, 1, generator=g), "acc_weight": 1.0 / 3 + 2.0 / 3}})
        allc = gather_class_code(mine)          # reduce=False: the merge itself runs on the device (GPU tests)
        q.put((rank, [(int(c["support_set_target"]), c["class_name"], c["class_code"]["acc_weight"],
                       c["class_code"]["cls_conv"].clone()) for c in allc]))
    finally:
        dist.destroy_process_group()


def test_gather_carries_accumulated_weights_in_double_precision():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_base, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, entries in out:
        assert [(e[0], e[1]) for e in entries] == [(7, "seven"), (7, "seven"), (3, "three")]
        assert entries[0][2] == 0.4 and entries[1][2] == 0.6 and entries[2][2] == 1.0 / 3 + 2.0 / 3   # exact doubles
        assert torch.equal(entries[0][3], torch.randn(1, 256, 1, 1, generator=torch.Generator().manual_seed(500)))


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


def _worker_losses(rank, world, port, q):
    """Training forward on 2 ranks: each rank holds its own query images; the positives and the
    centre-ness target sum are all-reduced (fcos_outputs.py:520-523, 557-558) -- the mirror's `_reduce_sum` /
    `_world_size` on a float64 pair, and the oracle's reduce hook, against a single-process recomputation."""
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        from oracle import upstream as up
        from oracle.meta_fcos_oracle import MetaFCOSOracle
        from sylph_few_shot_detection_b200 import modeling as M
        from sylph_few_shot_detection_b200 import weights as W
        from tests.cases import cfg_for
        cfg = cfg_for("COCO-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml", ["MODEL.PROPOSAL_GENERATOR.FREEZE_BBOX_BRANCH", False,
                                                                           "MODEL.PROPOSAL_GENERATOR.FREEZE", False])
        orc = MetaFCOSOracle(cfg, {"pixel_mean": torch.zeros(3, 1, 1), "pixel_std": torch.ones(3, 1, 1),
                                   **{k: v for k, v in W.synthetic_state_dict(cfg, 1).items() if k.startswith("backbone.")}})
        g = torch.Generator().manual_seed(40 + rank)
        sizes = [(12, 16), (6, 8), (3, 4), (2, 2), (1, 1)]
        logits = [torch.randn(1, 2, h, w, generator=g) - 2 for h, w in sizes]
        regs = [torch.rand(1, 4, h, w, generator=g) * 4 for h, w in sizes]
        ctrs = [torch.randn(1, 1, h, w, generator=g) for h, w in sizes]
        gts = [(torch.tensor([[8.0, 8.0, 90.0, 70.0], [30.0, 20.0, 60.0, 50.0]]) + 10 * rank, torch.tensor([3, 5]))]
        losses, ex = orc.fcos_losses(logits, regs, ctrs, gts, [3, 5], world_size=up.get_world_size(), reduce=up.reduce_sum)
        local, _ = orc.fcos_losses(logits, regs, ctrs, gts, [3, 5])            # world 1: local normalisers
        pair = torch.tensor([float(ex["num_pos"]), float(ex["ctr_targets_sum"])], dtype=torch.float64)
        total = M._reduce_sum(pair)
        q.put((rank, M._world_size(), {k: float(v) for k, v in losses.items()}, {k: float(v) for k, v in local.items()},
               pair.tolist(), total.tolist()))
    finally:
        dist.destroy_process_group()


def test_training_loss_normalisers_are_reduced_over_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_losses, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pos = [o[4][0] for o in out]
    ctr = [o[4][1] for o in out]
    assert all(p > 0 for p in pos) and pos[0] != pos[1] or ctr[0] != ctr[1]
    for rank, world, losses, local, pair, total in out:
        assert world == 2
        assert total[0] == pos[0] + pos[1] and abs(total[1] - (ctr[0] + ctr[1])) < 1e-9
        avg_pos, avg_ctr = max(total[0] / 2, 1.0), max(total[1] / 2, 1e-6)
        # same sums, other normalisers: loss_world2 = loss_local * local_norm / global_norm
        assert abs(losses["loss_fcos_cls"] - local["loss_fcos_cls"] * max(pair[0], 1.0) / avg_pos) < 1e-5 * abs(losses["loss_fcos_cls"])
        assert abs(losses["loss_fcos_ctr"] - local["loss_fcos_ctr"] * max(pair[0], 1.0) / avg_pos) < 1e-5 * abs(losses["loss_fcos_ctr"])
        assert abs(losses["loss_fcos_loc"] - local["loss_fcos_loc"] * max(pair[1], 1e-6) / avg_ctr) < 1e-5 * abs(losses["loss_fcos_loc"])


class _StandInExchangeEngine:
    """Host-side stand-in for Engine.exchange_setup / normalize_codes_exchange (the real one stores rows into peer GPU
    memory, tests/test_gpu_zexchange.py): rows x 2 as the "normalisation", delivery through a gloo all-reduce of
    disjoint row ranges.  What is under test is `exchange_codes_peer`'s shard arithmetic and output schema."""
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def exchange_setup(self, group, max_classes):
        self.max_classes = max_classes

    def normalize_codes_exchange(self, raw, class_offset, n_total):
        self.calls.append((0 if raw is None else raw.shape[0], class_offset, n_total))
        rows = torch.zeros((n_total, 257))
        if raw is not None:
            rows[class_offset:class_offset + raw.shape[0]] = raw * 2
        dist.all_reduce(rows)
        return rows


def _worker_peer(rank, world, port, n_classes, q):
    import types

    from sylph_few_shot_detection_b200.runner import exchange_codes_peer
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        mine = []
        for c in shard_range(n_classes, world, rank):
            g = torch.Generator().manual_seed(100 + c)
            mine.append({"support_set_target": torch.tensor(c), "class_name": f"class{c}",
                         "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g),
                                        "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}})
        counts = [len(shard_range(n_classes, world, r)) for r in range(world)]
        meta = [(torch.tensor(c), f"class{c}") for c in range(n_classes)]
        eng = _StandInExchangeEngine()
        allc = exchange_codes_peer(types.SimpleNamespace(engine=eng), mine, counts, meta)
        packed = format_class_codes_shared(allc)
        q.put((rank, eng.calls, [int(c["support_set_target"]) for c in allc], [c["class_name"] for c in allc],
               [tuple(c["class_code"]["cls_bias"].shape) for c in allc], packed["cls_conv"].clone(), packed["cls_bias"].clone()))
    finally:
        dist.destroy_process_group()


def test_exchange_codes_peer_world2_shards_offsets_and_schema():
    n = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_peer, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gens = [torch.Generator().manual_seed(100 + c) for c in range(n)]
    conv = torch.cat([torch.randn(1, 256, 1, 1, generator=g) for g in gens]) * 2
    bias = torch.cat([torch.randn(1, 1, 1, 1, generator=g).reshape(1) for g in gens]) * 2
    assert out[0][1] == [(3, 0, 5)] and out[1][1] == [(2, 3, 5)]            # (n_local, class_offset, n_total) per rank
    for rank, calls, ids, names, bias_shapes, got_conv, got_bias in out:
        assert ids == list(range(n)) and names == [f"class{c}" for c in range(n)]
        assert bias_shapes == [(1,)] * n                                   # process_bias output shape, code_generator.py:853
        assert torch.equal(got_conv, conv) and torch.equal(got_bias, bias)


class _RecordingLib:
    """Stands in for libsylph_b200.so in `Engine.exchange_setup`: hands out a recognisable 64-byte handle per rank and
    records what `sylph_exchange_connect` receives (the real calls need a GPU: tests/test_gpu_zexchange.py)."""

    def __init__(self):
        self.created, self.connected, self.destroyed = None, None, 0

    def sylph_exchange_create(self, h, world, rank, max_classes, handle):
        self.created = (world, rank, max_classes)
        for i in range(64):
            handle[i] = (rank * 64 + i) % 251
        return 0

    def sylph_exchange_connect(self, h, handles):
        self.connected = None if handles is None else bytes(handles)
        return 0

    def sylph_exchange_destroy(self, h):
        self.destroyed += 1

    def sylph_last_error(self, h):
        return b""


def _worker_handles(rank, world, port, q):
    from sylph_few_shot_detection_b200.runtime import Engine
    from tests.cases import init_gloo
    init_gloo(rank, world, port)          # `port` is a FileStore path (tests/cases.fresh_rendezvous)
    try:
        eng = Engine.__new__(Engine)            # no device here: only the handle exchange of exchange_setup is under test
        eng.lib, eng.h, eng.device = _RecordingLib(), 1, torch.device("cpu")
        eng.exchange_setup(None, max_classes=77)
        first = (eng.lib.created, eng.lib.connected)
        eng.exchange_setup(None, max_classes=77)            # idempotent for the same geometry
        assert eng.lib.created == first[0] and eng.lib.destroyed == 0
        # (exchange_teardown synchronises a CUDA device first: exercised by the GPU tests only)
        q.put((rank, first[0], first[1]))
        eng.h = None                                         # nothing for __del__ to destroy
    finally:
        dist.destroy_process_group()


def test_exchange_setup_swaps_the_ipc_handles_in_rank_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_handles, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes((r * 64 + i) % 251 for r in range(2) for i in range(64))
    for rank, created, connected in out:
        assert created == (2, rank, 77)
        assert connected == want and len(connected) == 128
