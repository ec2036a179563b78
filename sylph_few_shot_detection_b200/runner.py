"""Meta-test orchestration: the callers of the hot path, restated for the B200 build.

Mirrors, step for step, `MetaFCOSRunner._do_test_meta_learning` (sylph/runner/meta_fcos_runner.py:451-672, SURVEY.md
section 3.2):
  B. inference_on_support_set_dataset ........ sylph/evaluation/meta_learn_evaluation.py:256-365
  C. _gather_class_code ...................... sylph/runner/meta_fcos_runner.py:381-439
  D. inference_normalization ................. sylph/evaluation/meta_learn_evaluation.py:105-116
  E. format_class_codes_shared ............... sylph/evaluation/meta_learn_evaluation.py:71-103
  F. inference_on_dataset_with_class_codes ... sylph/evaluation/meta_learn_evaluation.py:367-470
The reference runs B and F as batch-1 Python loops with a device synchronise per item; here all classes of a rank
go through one backbone batch and one code-generation launch sequence, all query images through one head launch
sequence, and the gather moves a fixed-stride (classes, 257) fp32 DEVICE buffer with one NCCL all-gather instead of
pickled CPU tensors through all_gather_object.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

import torch
import torch.distributed as dist

from .config import CfgNode, get_default_cfg
from .modeling import META_ARCH_REGISTRY, MetaOneStageDetector, build_model

CODE_STRIDE = 257


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous, balanced shards -- detectron2 `InferenceSampler._get_local_indices` (recent versions): the first
    `n % world` ranks get one extra item.  (sylph/data/build.py:578-592 builds the support loader with it.)"""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return range(begin, begin + base + (1 if rank < extra else 0))


def balanced_query_assignment(support_images_per_rank: Sequence[int], n_query: int, query_cost: float = 1.7) -> List[List[int]]:
    """Query image -> rank assignment for ONE episode whose classes are already sharded: every query image goes, in
    order, to the rank with the smallest load so far (support images + `query_cost` image-equivalents per query image it
    already holds: a query image pays the backbone AND the FCOS towers); ties go to the lowest rank.  Deterministic,
    computed identically on every rank.  The reference shards query images contiguously with `InferenceSampler`
    (sylph/data/build.py:749-755) whatever the class shards cost; with 20 classes on 8 GPUs that leaves the ranks that
    hold 3 classes (15 support images) with a query image on top while the ranks with 2 classes wait in the all-gather."""
    load = [float(v) for v in support_images_per_rank]
    out: List[List[int]] = [[] for _ in load]
    for q in range(n_query):
        r = min(range(len(load)), key=lambda i: (load[i], i))
        out[r].append(q)
        load[r] += query_cost
    return out


def inference_on_support_set(model: MetaOneStageDetector, support_items: Sequence[Dict[str, Any]],
                             features_in_slot: bool = False) -> List[Dict]:
    """Step B for this rank's classes.  Each item: {"support_set": [K records], "support_set_target", "class_name"}.
    Returns the reference's list schema (meta_learn_evaluation.py:305-326) with RAW codes on the device."""
    if len(support_items) == 0:
        return []
    with torch.no_grad():
        codes = model.forward_class_codes_batched(list(support_items), features_in_slot=features_in_slot)
    out = []
    for item, code in zip(support_items, codes):
        out.append({"support_set_target": item["support_set_target"], "class_name": item.get("class_name", ""),
                    "class_code": code})
    return out


def _target_id(t) -> int:
    return int(t.item()) if torch.is_tensor(t) else int(t)


def inference_on_support_set_base(model: MetaOneStageDetector, chunk_items: Sequence[Dict[str, Any]],
                                  chunks_per_batch: int = 4) -> List[Dict]:
    """Base-class "all ground truths" path for this rank's chunks (inference_on_support_set_dataset_base,
    sylph/evaluation/meta_learn_evaluation.py:118-254).  Every item is ONE chunk of <= 10 support records of one class
    and carries "len" / "total_len" (sylph/data/.../meta_lvis.py:303-306); its code is weighted by len / total_len and
    accumulated per class in arrival order.  Returns the reference's per-rank list: one entry per class id seen here,
    `class_code` = {"cls_conv", "cls_bias", "acc_weight"}; codes stay on the device, `acc_weight` is a Python float
    accumulated in double precision like the reference's."""
    if len(chunk_items) == 0:
        return []
    eng = model.engine
    cids = [_target_id(it["support_set_target"]) for it in chunk_items]
    weights = [float(it["len"]) / it["total_len"] for it in chunk_items]
    order: List[int] = []           # class ids in first-appearance order (dict insertion order in the reference)
    names: Dict[int, str] = {}
    acc_w: Dict[int, float] = {}
    for cid, w, it in zip(cids, weights, chunk_items):
        if cid not in acc_w:
            order.append(cid)
            acc_w[cid] = 0.0
        acc_w[cid] += w
        names[cid] = it.get("class_name", "")
    slot_of = {cid: i for i, cid in enumerate(order)}
    acc = torch.zeros((len(order), CODE_STRIDE), device=eng.device, dtype=torch.float32)
    with torch.no_grad():
        for i in range(0, len(chunk_items), chunks_per_batch):
            batch = list(chunk_items[i:i + chunks_per_batch])
            codes = model.forward_class_codes_batched(batch)
            rows = torch.cat([torch.cat([c["cls_conv"].reshape(1, 256), c["cls_bias"].reshape(1, 1)], dim=1) for c in codes])
            eng.accumulate_codes(rows, [slot_of[c] for c in cids[i:i + chunks_per_batch]], weights[i:i + chunks_per_batch], acc)
    return [{"support_set_target": cid, "class_name": names[cid],
             "class_code": {"cls_conv": acc[slot_of[cid], :256].reshape(1, 256, 1, 1),
                            "cls_bias": acc[slot_of[cid], 256:].reshape(1, 1, 1, 1), "acc_weight": acc_w[cid]}}
            for cid in order]


def reduce_class_code(out_codes: List[Dict], engine=None) -> List[Dict]:
    """`reduce_class_code` (sylph/modeling/code_generator/utils.py:397-427): merge the entries of one class id (partial
    sums from different ranks) by adding them in list order, divide by the accumulated weight where it is not 1
    (|1 - acc_weight| > 1e-6) and drop `acc_weight`.  Output order = first appearance of each class id.  The sums run
    in one kernel over a dense (occurrence, class, 257) device buffer (`sylph_reduce_codes`)."""
    if len(out_codes) == 0:
        return out_codes
    assert "class_code" in out_codes[0]
    order: List[int] = []
    occ: Dict[int, List[Dict]] = {}
    other: Dict[int, Dict] = {}
    for c in out_codes:
        cid = _target_id(c["support_set_target"])
        assert "class_code" in c
        if cid not in occ:
            order.append(cid)
            occ[cid] = []
            other[cid] = {k: v for k, v in c.items() if k != "class_code"}
        occ[cid].append(c["class_code"])
    dev = None
    for c in out_codes:
        if c["class_code"]["cls_conv"].is_cuda:
            dev = c["class_code"]["cls_conv"].device
            break
    if engine is None or dev is None:
        raise RuntimeError("reduce_class_code runs on the device: pass the model's engine and device-resident codes "
                           "(the B200 path has no CPU fallback)")
    n_parts = max(len(v) for v in occ.values())
    parts = torch.zeros((n_parts, len(order), CODE_STRIDE), device=dev, dtype=torch.float32)
    divisor = []
    for i, cid in enumerate(order):
        acc_weight = 0
        for j, code in enumerate(occ[cid]):
            parts[j, i, :256] = code["cls_conv"].reshape(-1).to(dev)
            parts[j, i, 256] = code["cls_bias"].reshape(-1)[0].to(dev)
            acc_weight = acc_weight + code["acc_weight"]          # functools.reduce(lambda x, y: x + y[key], lst, 0)
        divisor.append(float(acc_weight) if abs(1.0 - acc_weight) > 1e-6 else 0.0)
    red = engine.reduce_codes(parts, divisor)
    results = []
    for i, cid in enumerate(order):
        r = dict(other[cid])
        r["class_code"] = {"cls_conv": red[i, :256].reshape(1, 256, 1, 1), "cls_bias": red[i, 256:].reshape(1, 1, 1, 1)}
        results.append(r)
    return results


def replace_class_code(support_set_class_code: List[Dict], target_class_codes: List[Dict], device) -> List[Dict]:
    """`replace_class_code` (sylph/modeling/code_generator/utils.py:376-394): few-shot codes whose class id also has
    an all-GT base code take that code (the first one of the id); everything ends up on `device`."""
    target: Dict[int, Dict] = {}
    for c in target_class_codes:
        target.setdefault(_target_id(c["support_set_target"]), c["class_code"])
    results = []
    for c in support_set_class_code:
        r = dict(c)
        cid = _target_id(c["support_set_target"])
        code = target[cid] if cid in target else c["class_code"]
        r["class_code"] = {k: (v.to(device) if torch.is_tensor(v) else torch.tensor(v).to(device)) for k, v in code.items()}
        results.append(r)
    return results


def _code_rows(codes: List[Dict], device) -> torch.Tensor:
    """(n, 257) rows from the list schema with three concatenations (no per-class kernels, no host sync)."""
    conv = torch.cat([c["class_code"]["cls_conv"].reshape(1, 256) for c in codes], dim=0).to(device, torch.float32)
    bias = torch.cat([c["class_code"]["cls_bias"].reshape(1, 1) for c in codes], dim=0).to(device, torch.float32)
    return torch.cat([conv, bias], dim=1)


def gather_class_code_known_shards(sub_class_codes: List[Dict], counts: Sequence[int], meta: Sequence[Tuple[Any, str]],
                                   group=None, device: Optional[torch.device] = None) -> List[Dict]:
    """Step C when every rank already knows the shard sizes and the (support_set_target, class_name) of every class --
    the case of `run_episode`, where the class list is global and only the CODES are produced per rank.  Exactly ONE
    collective (all_gather_into_tensor of a fixed-stride [max_shard, 257] fp32 device buffer) and no host
    synchronisation: the returned dicts hold views of the gathered buffer, the stream keeps running ahead.
    Same result as `gather_class_code` (tests/test_runner_dist.py)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and len(sub_class_codes) == counts[rank] and len(meta) == sum(counts)
    if device is None:
        device = (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
                  else torch.device("cpu"))
    max_n = max(max(counts), 1)
    if len(sub_class_codes) == max_n:
        buf = _code_rows(sub_class_codes, device).contiguous()
    else:
        buf = torch.zeros((max_n, CODE_STRIDE), dtype=torch.float32, device=device)
        if len(sub_class_codes):
            buf[:len(sub_class_codes)] = _code_rows(sub_class_codes, device)
    gathered = torch.empty((world * max_n, CODE_STRIDE), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    out, k = [], 0
    for r in range(world):
        for i in range(counts[r]):
            row = gathered[r * max_n + i]
            target, name = meta[k]
            k += 1
            out.append({"support_set_target": target, "class_name": name,
                        "class_code": {"cls_conv": row[:256].reshape(1, 256, 1, 1), "cls_bias": row[256:257].reshape(1, 1, 1, 1)}})
    return out


def gather_class_code(sub_class_codes: List[Dict], group=None, device: Optional[torch.device] = None,
                      reduce: bool = False, engine=None) -> List[Dict]:
    """Step C: all ranks end up with the codes of ALL classes, ordered by rank then local order (what concatenating
    the all_gather_object output gives, meta_fcos_runner.py:386-396).  One all-gather of a padded
    [max_shard, 258] fp32 buffer (257 code floats + class id; class names travel separately only if present) and,
    for the base-class path, one of the float64 accumulated weights.  `reduce=True` then merges the entries of a
    class id with `reduce_class_code` (:432-439).  World size 1 short-circuits like the reference (:430-431)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return reduce_class_code(sub_class_codes, engine) if reduce else sub_class_codes
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(sub_class_codes)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in counts]
    max_n = max(max(counts), 1)
    buf = torch.zeros((max_n, CODE_STRIDE + 1), dtype=torch.float32, device=device)
    wbuf = torch.full((max_n,), -1.0, dtype=torch.float64, device=device)
    for i, c in enumerate(sub_class_codes):
        buf[i, :256] = c["class_code"]["cls_conv"].reshape(-1).to(device)
        buf[i, 256] = c["class_code"]["cls_bias"].reshape(-1)[0].to(device)
        buf[i, 257] = float(_target_id(c["support_set_target"]))
        if "acc_weight" in c["class_code"]:
            wbuf[i] = float(c["class_code"]["acc_weight"])
    gathered = torch.empty((world * max_n, CODE_STRIDE + 1), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    wall = torch.empty((world * max_n,), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(wall, wbuf, group=group)
    wall = wall.cpu().tolist()
    names = [None] * world
    dist.all_gather_object(names, [c.get("class_name", "") for c in sub_class_codes], group=group)
    out = []
    for r in range(world):
        for i in range(counts[r]):
            row = gathered[r * max_n + i]
            code = {"cls_conv": row[:256].reshape(1, 256, 1, 1).clone(), "cls_bias": row[256:257].reshape(1, 1, 1, 1).clone()}
            if wall[r * max_n + i] >= 0:
                code["acc_weight"] = wall[r * max_n + i]
            out.append({"support_set_target": torch.tensor(int(row[257].item())), "class_name": names[r][i], "class_code": code})
    return reduce_class_code(out, engine) if reduce else out


def inference_normalization(model: MetaOneStageDetector, all_class_codes: List[Dict]) -> List[Dict]:
    """Step D (meta_learn_evaluation.py:105-116)."""
    with torch.no_grad():
        return model(None, class_code=all_class_codes, run_type="meta_learn_normalize_code")


def format_class_codes_shared(all_class_codes: List[Dict], device=None) -> Dict[str, torch.Tensor]:
    """Step E (meta_learn_evaluation.py:71-103): index by support_set_target, concatenate, flatten the bias."""
    n = len(all_class_codes)
    if n == 0:
        return all_class_codes          # the reference returns the (empty) list itself (:86-87)
    slots: List[Optional[Dict]] = [None] * n
    for c in all_class_codes:
        i = _target_id(c["support_set_target"])
        if not 0 <= i < n:        # the reference indexes a list of n entries with the target id: IndexError
            raise IndexError(f"support_set_target {i} outside [0, {n}): class ids must be 0..n-1 (meta_learn_evaluation.py:93-96)")
        slots[i] = c["class_code"]
    missing = [i for i, v in enumerate(slots) if v is None]
    if missing:                   # ... and concatenates the list: a hole (duplicate / missing id) is a TypeError there
        raise TypeError(f"no class code for class id(s) {missing[:8]}: ids must be exactly 0..{n - 1} without duplicates")
    conv = torch.cat([v["cls_conv"] for v in slots], dim=0)
    bias = torch.cat([v["cls_bias"].reshape(-1) for v in slots], dim=0)
    if device is not None:
        conv, bias = conv.to(device), bias.to(device)
    return {"cls_conv": conv, "cls_bias": bias}


def inference_with_class_codes(model: MetaOneStageDetector, query_items: Sequence[Dict[str, Any]],
                               class_codes: Dict[str, torch.Tensor], batch_size: int = 16,
                               features_in_slot: bool = False) -> List[Dict]:
    """Step F for this rank's query images: [{"instances": Instances}] per image, in input order."""
    out: List[Dict] = []
    with torch.no_grad():
        if features_in_slot:
            assert len(query_items) <= batch_size
            return model.forward_instances(list(query_items), class_codes, features_in_slot=True)
        for i in range(0, len(query_items), batch_size):
            out.extend(model(list(query_items[i:i + batch_size]), class_code=class_codes, run_type="meta_learn_test_instance"))
    return out


_COPY_STREAMS: Dict[Any, Any] = {}


def _copy_stream(device):
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def query_indices_of_rank(support_items: Sequence[Dict[str, Any]], n_query: int, world: int, rank: int,
                          balance_queries: bool = False) -> List[int]:
    """Indices (into the episode's query list) of the query images rank `rank` detects on in `run_episode`."""
    if not balance_queries:
        return list(shard_range(n_query, world, rank))
    per_rank = [sum(len(support_items[i]["support_set"]) for i in shard_range(len(support_items), world, r)) for r in range(world)]
    return balanced_query_assignment(per_rank, n_query)[rank]


def exchange_codes_peer(model: MetaOneStageDetector, sub_class_codes: List[Dict], counts: Sequence[int],
                        meta: Sequence[Tuple[Any, str]], group=None) -> List[Dict]:
    """Steps C + D of the sharded episode in one fused device step: this rank's raw codes are normalised and stored
    straight into every rank's exchange buffer over NVLink (`Engine.normalize_codes_exchange`), replacing
    `_gather_class_code` (meta_fcos_runner.py:381-396) + `inference_normalization` (meta_learn_evaluation.py:105-116).
    Returns the NORMALISED codes of all classes in the list schema (views of one (classes, 257) device buffer); same
    values as `gather_class_code_known_shards` followed by `inference_normalization`."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    assert len(counts) == world and len(sub_class_codes) == counts[rank] and len(meta) == sum(counts)
    engine = model.engine
    engine.exchange_setup(group, max_classes=max(2048, sum(counts)))
    raw = _code_rows(sub_class_codes, engine.device) if sub_class_codes else None
    rows = engine.normalize_codes_exchange(raw, sum(counts[:rank]), sum(counts))
    return [{"support_set_target": target, "class_name": name,
             "class_code": {"cls_conv": rows[k, :256].reshape(1, 256, 1, 1), "cls_bias": rows[k, 256:257].reshape(1)}}
            for k, (target, name) in enumerate(meta)]


def code_exchange_mode() -> str:
    """"nccl" (one all_gather_into_tensor, then normalisation on every rank), "peer" (normalisation fused with the
    all-gather over NVLink peer memory) or "auto" (default): peer when every rank of the group could map every other rank's
    exchange buffer (GPUs of one box with peer access), the collective otherwise.  Measured on 8 B200s (profiles/r02_bench_n8.json,
    r02_cfg5_*_n8.json): the exchange alone 0.21 ms (peer) against 0.49 ms (all-gather + normalisation of all classes), the
    1203-class sweep 1.70 against 1.98 ms, the 20-way episode equal within noise.  SYLPH_CODE_EXCHANGE overrides."""
    return os.environ.get("SYLPH_CODE_EXCHANGE", "auto")


def _resolve_exchange(model: MetaOneStageDetector, mode: Optional[str], n_classes: int, group=None) -> str:
    """"auto" -> "peer" / "nccl", decided ONCE per engine and agreed by all ranks (a MIN all-reduce of the local outcome of the
    IPC set-up), so no rank can end up waiting in a collective the others do not enter."""
    mode = mode or code_exchange_mode()
    if mode != "auto":
        return mode
    engine = model.engine
    cached = getattr(engine, "_auto_exchange", None)
    if cached is not None:
        return cached
    decided = "nccl"
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl" and hasattr(engine, "exchange_setup"):
        ok = 1
        try:
            engine.exchange_setup(group, max_classes=max(2048, n_classes))
        except RuntimeError:
            ok = 0
        flag = torch.tensor([ok], device=engine.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 1:
            decided = "peer"
        elif ok:
            engine.exchange_teardown(group)
    engine._auto_exchange = decided
    return decided


def run_episode(model: MetaOneStageDetector, support_items: Sequence[Dict[str, Any]], query_items: Sequence[Dict[str, Any]],
                group=None, shard: bool = True, return_device: bool = False, balance_queries: bool = False,
                exchange: Optional[str] = None) -> List[Dict]:
    """One meta-test episode (steps B-F).  With a process group and `shard=True`, classes and query images are split
    contiguously over the ranks (InferenceSampler semantics) with ONE exchange step in between; each rank returns the
    detections of its own query shard.  `balance_queries=True` hands the query images to the least-loaded ranks instead
    (`balanced_query_assignment`; the rank's query indices are `query_indices_of_rank(...)`).  `exchange`: "nccl" = one
    all_gather_into_tensor of the raw codes, "peer" = `exchange_codes_peer` (default: `code_exchange_mode()`)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if world > 1 and shard:
        my_support = [support_items[i] for i in shard_range(len(support_items), world, rank)]
        my_query = [query_items[i] for i in query_indices_of_rank(support_items, len(query_items), world, rank, balance_queries)]
    else:
        my_support, my_query = list(support_items), list(query_items)
    # host-resident query images start their H2D copy on a side stream now, so it overlaps the support pass
    ready = None
    if torch.cuda.is_available() and any(not q["image"].is_cuda for q in my_query):
        side = _copy_stream(model.device)
        with torch.cuda.stream(side):
            staged = [dict(q, image=q["image"].to(model.device, non_blocking=True)) for q in my_query]
        ready = torch.cuda.Event()
        ready.record(side)
        my_query = staged
    # support and query images already on the device: ONE bottom-up trunk pass for both batches (the FPN of each
    # lands in its own slot); the engine falls back to two passes when the batches pad to different sizes
    merged = False
    if ready is None and my_support and 0 < len(my_query) <= 16 and torch.cuda.is_available():
        sup_imgs = [r["image"] for it in my_support for r in it["support_set"]]
        qry_imgs = [q["image"] for q in my_query]
        # one shared support batch only when every class pads to the same size (each reference call pads to its own maximum)
        one_size = len({model.class_padded_size(it) for it in my_support}) == 1
        if one_size and all(t.is_cuda for t in sup_imgs + qry_imgs) and len(sup_imgs) + len(qry_imgs) <= 64:
            from .runtime import SLOT_QUERY, SLOT_SUPPORT
            model.engine.extract_features_multi([(SLOT_SUPPORT, sup_imgs), (SLOT_QUERY, qry_imgs)])
            merged = True
    sub_codes = inference_on_support_set(model, my_support, features_in_slot=merged)
    if world > 1 and shard:
        # the class list is global: shard sizes, ids and names are known everywhere, only the codes travel
        counts = [len(shard_range(len(support_items), world, r)) for r in range(world)]
        meta = [(it["support_set_target"], it.get("class_name", "")) for it in support_items]
        exchange = _resolve_exchange(model, exchange, len(support_items), group)
        if exchange == "peer":
            all_codes = exchange_codes_peer(model, sub_codes, counts, meta, group=group)    # already normalised
        else:
            all_codes = inference_normalization(model, gather_class_code_known_shards(sub_codes, counts, meta, group=group))
    else:
        all_codes = inference_normalization(model, sub_codes)
    packed = format_class_codes_shared(all_codes, device=model.device)
    if ready is not None:
        torch.cuda.current_stream().wait_event(ready)
    if return_device:
        if not my_query:            # a rank without query images (more ranks than images, or balance_queries)
            return None, None, []
        assert len(my_query) <= 16
        with torch.no_grad():
            return model.forward_instances_device(list(my_query), packed, features_in_slot=merged)
    results = inference_with_class_codes(model, my_query, packed, features_in_slot=merged)
    if world > 1 and shard and exchange == "peer" and my_query:
        model.engine.exchange_poll()   # the results above are on the host, so the flag of this episode is too: incomplete codes raise
    return results


class EpisodeFuture:
    """Detections of an episode enqueued with `EpisodePipeline.run_async`: `result()` waits for the device-to-host copy
    and returns the reference's schema `[{"instances": Instances}]` with HOST tensors (what the evaluators consume,
    sylph/evaluation/meta_learn_evaluation.py:430-437).  At most 4 futures may be outstanding (pinned ring)."""

    def __init__(self, dets_host: torch.Tensor, counts_host: torch.Tensor, out_sizes, done):
        self._dets, self._counts, self._out_sizes, self._done = dets_host, counts_host, out_sizes, done

    def result(self) -> List[Dict]:
        from .modeling import instances_from_detections
        self._done.synchronize()
        counts = self._counts.tolist()
        dets = self._dets.clone()    # the pinned ring slot is reused by a later episode
        return [{"instances": r} for r in instances_from_detections(dets, counts, self._out_sizes)]


class EpisodePipeline:
    """Throughput mode for a stream of episodes with HOST-resident inputs: `submit` starts the host-to-device copies
    of an episode on a side stream, `run` executes it on the current stream once its copies have landed.  Submitting
    episode i+1 before running episode i overlaps its H2D traffic with the compute of episode i.

        h = pipe.submit(s0, q0)
        for next_s, next_q in episodes:
            nxt = pipe.submit(next_s, next_q)
            results = pipe.run(h)
            h = nxt
    """

    def __init__(self, model: MetaOneStageDetector):
        self.model = model
        self._ring: List[Any] = []   # pinned (dets, counts) host buffers of the last 4 asynchronous episodes
        self._ring_pos = 0

    def submit(self, support_items: Sequence[Dict[str, Any]], query_items: Sequence[Dict[str, Any]]):
        dev = self.model.device
        side = _copy_stream(dev)
        with torch.cuda.stream(side):
            sup = []
            for item in support_items:
                recs = [dict(r, image=r["image"].to(dev, non_blocking=True)) for r in item["support_set"]]
                sup.append(dict(item, support_set=recs))
            qry = [dict(q, image=q["image"].to(dev, non_blocking=True)) for q in query_items]
        ev = torch.cuda.Event()
        ev.record(side)
        return sup, qry, ev

    def run_async(self, handle) -> "EpisodeFuture":
        """Enqueue the episode and an asynchronous copy of its detections into pinned host memory; nothing is
        synchronised, so the caller can submit / enqueue the NEXT episode before it reads this one's results
        (`future.result()`): the device never waits for the host between episodes."""
        sup, qry, ev = handle
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        dets, counts, out_sizes = run_episode(self.model, sup, qry, shard=False, return_device=True)
        def fresh():
            return (torch.empty(dets.shape, dtype=dets.dtype, pin_memory=True),
                    torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True))
        k = self._ring_pos % 4
        self._ring_pos += 1
        if k >= len(self._ring):
            self._ring.append(fresh())            # the first four episodes allocate the ring
        elif self._ring[k][0].shape != dets.shape:
            self._ring[k] = fresh()               # another batch size: replace the slot (its old future was 4 episodes ago)
        slot = self._ring[k]
        slot[0].copy_(dets, non_blocking=True)
        slot[1].copy_(counts, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        for item in sup:
            for r in item["support_set"]:
                r["image"].record_stream(cur)
        for q in qry:
            q["image"].record_stream(cur)
        return EpisodeFuture(slot[0], slot[1], out_sizes, done)

    def run(self, handle) -> List[Dict]:
        sup, qry, ev = handle
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        out = run_episode(self.model, sup, qry, shard=False)
        # the staged tensors were allocated on the copy stream: tell the caching allocator they are in use on the
        # compute stream, otherwise the next submit() could recycle their memory while kernels still read it
        for item in sup:
            for r in item["support_set"]:
                r["image"].record_stream(cur)
        for q in qry:
            q["image"].record_stream(cur)
        return out


class EpisodeGraph:
    """One meta-test episode of FIXED shape (classes x shots, query count, image size) captured into a CUDA graph: the
    ~125 kernel launches of support backbone + code generation + normalisation + query backbone + head + proposals
    replay as ONE graph launch, programmatic-dependent-launch edges included.  Inputs live in static device buffers:

        g = EpisodeGraph(model, n_way=5, n_shot=5, n_query=8, image_hw=(800, 1333))
        g.support[i].copy_(img); g.query[j].copy_(img); g.boxes.copy_(boxes_xyxy)     # any stream-ordered writes
        dets, counts = g.replay()          # (n_query, max_dets, 9) detections and per-image counts, device tensors

    Results are bit-identical to the eager calls (same kernels, same order; tests/test_gpu_cases.py)."""

    def __init__(self, model: MetaOneStageDetector, n_way: int, n_shot: int, n_query: int, image_hw, warmup: int = 2):
        from .runtime import SLOT_QUERY, SLOT_SUPPORT
        self.model, self.eng = model, model.engine
        dev = model.device
        h, w = image_hw
        self.support = [torch.zeros((3, h, w), dtype=torch.uint8, device=dev) for _ in range(n_way * n_shot)]
        self.query = [torch.zeros((3, h, w), dtype=torch.uint8, device=dev) for _ in range(n_query)]
        self.boxes = torch.tensor([[0.25 * w, 0.25 * h, 0.75 * w, 0.75 * h]] * (n_way * n_shot), dtype=torch.float32, device=dev)
        self._roi_image = list(range(n_way * n_shot))
        self._offsets = list(range(0, n_way * n_shot + 1, n_shot))
        self._slots = (SLOT_SUPPORT, SLOT_QUERY)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # allocations, plane tables and kernel attributes settle before capture
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.dets, self.counts = self._run()

    def _run(self):
        sup_slot, qry_slot = self._slots
        self.eng.extract_features_multi([(sup_slot, self.support), (qry_slot, self.query)])
        out, self.codes = self.eng.generate_and_detect(
            sup_slot, qry_slot, self.boxes, self._roi_image, self._offsets,
            normalize=not isinstance(self.model.code_generator, _roi_encoder_type()))
        return out

    def replay(self):
        self.graph.replay()
        return self.dets, self.counts


def _is_main_process() -> bool:
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def _roi_encoder_type():
    from .modeling import ROIEncoder
    return ROIEncoder


class MetaFCOSRunner:
    """The d2go-runner surface the reference CLI drives (`create_runner("sylph.runner.MetaFCOSRunner")`,
    sylph/runner/meta_fcos_runner.py:92-114, 381-382, 674-701), reduced to the inference hot path."""

    def get_default_cfg(self) -> CfgNode:
        return get_default_cfg()

    def build_model(self, cfg, eval_only: bool = True) -> MetaOneStageDetector:
        if not eval_only:
            raise NotImplementedError("training is outside the B200 inference path")
        model = build_model(cfg)
        type(self)._model = model
        return model

    _model = None   # reduce_class_code sums on the device through the last built model's C-ABI context

    @classmethod
    def _gather_class_code(cls, sub_class_codes: List[Dict], reduce: bool = False) -> List[Dict]:
        """meta_fcos_runner.py:381-439; `reduce=True` is the base-class all-GT path (reduce_class_code)."""
        return gather_class_code(sub_class_codes, reduce=reduce, engine=cls._model.engine if cls._model is not None else None)

    # ---- loaders and evaluators are the caller's (datasets, mappers and COCO / LVIS evaluators are outside the hot path):
    # subclass and override, exactly the methods the reference runner defines (meta_fcos_runner.py:163-230, 232-300).
    def build_episodic_learning_detection_test_support_set_loader(self, cfg, dataset_name: str, seed: int):
        """Iterable of batches `[item]` (batch size 1): item = {"support_set": [K records], "support_set_target", "class_name"}."""
        raise NotImplementedError("dataset loaders are outside the B200 hot path: override this builder (see INTEGRATION.md)")

    def build_episodic_learning_detection_test_support_set_base_loader(self, cfg, dataset_name: str):
        """Iterable of batches `[chunk]`: the items above plus "len" / "total_len" (<= 10 boxes per chunk, meta_lvis.py:303-306)."""
        raise NotImplementedError("dataset loaders are outside the B200 hot path: override this builder (see INTEGRATION.md)")

    def build_episodic_learning_detection_test_query_loader(self, cfg, dataset_name: str):
        """Iterable of batches of query records {"image", "height", "width"[, "instances"]}."""
        raise NotImplementedError("dataset loaders are outside the B200 hot path: override this builder (see INTEGRATION.md)")

    def get_evaluator(self, cfg, dataset_name: str, output_folder: Optional[str] = None):
        """An object with reset() / process(inputs, outputs) / evaluate() (detectron2 DatasetEvaluator protocol)."""
        raise NotImplementedError("evaluators are outside the B200 hot path: override get_evaluator (see INTEGRATION.md)")

    def thing_classes(self, cfg, dataset_name: str) -> Optional[Sequence[str]]:
        """`MetadataCatalog.get(dataset_name).thing_classes` (:501); None skips the class-count assert (:540-542)."""
        return None

    def _do_test_meta_learning(self, cfg, model, train_iter=None, model_tag: str = "default"):
        """`MetaFCOSRunner._do_test_meta_learning` (meta_fcos_runner.py:451-672), the evaluation sequence of the reference:
        per test repetition (`TEST.REPEAT_TEST` at the final iteration) and dataset -- class codes from the support-set
        loader (B), gather (C), optionally the base-class all-GT codes (gather with reduce, `replace_class_code`),
        normalisation (D), packing (E), detection over the query loader into the evaluator (F).  Steps B and F go
        through the batched entry points (the rank's classes in shared backbone batches, grouped by padded size and bounded
        per pass; query batches as they come), everything else is the reference's own sequence.  On the main process:
        {"seed<k>": {dataset: evaluator results}, model_tag: {dataset: bbox metrics averaged over the repetitions, plus
        AP*_avg / AP*_std at the final iteration}} (meta_fcos_runner.py:602-634)."""
        from collections import OrderedDict
        assert len(cfg.DATASETS.TEST)
        max_iter = (cfg.get("SOLVER") or {}).get("MAX_ITER")          # solver keys are not part of the inference config tree
        is_final = (train_iter is None) or (max_iter is not None and train_iter == max_iter - 1)
        type(self)._model = model
        results = OrderedDict()
        results[model_tag] = OrderedDict()
        num_repeat_test = cfg.TEST.REPEAT_TEST if is_final else 1
        for seed in range(num_repeat_test):
            results[f"seed{seed}"] = OrderedDict()
            for dataset_name in cfg.DATASETS.TEST:
                if "base" in dataset_name and cfg.MODEL.META_LEARN.EVAL_WITH_PRETRAINED_CODE:
                    raise NotImplementedError("inference with the pretrained class logits (base detector) is not on the B200 path")
                output_folder = None
                if cfg.get("OUTPUT_DIR", ""):
                    output_folder = os.path.join(cfg.OUTPUT_DIR, "inference", model_tag,
                                                 str(train_iter) if train_iter is not None else "final", dataset_name, str(seed))
                support_loader = self.build_episodic_learning_detection_test_support_set_loader(cfg, dataset_name, seed)
                items = []
                for inputs in support_loader:
                    assert len(inputs) == 1, "inputs' batch size is not 1"          # meta_learn_evaluation.py:299
                    items.append(inputs[0])
                sub_class_codes = inference_on_support_set(model, items)
                if output_folder is not None:                                        # :316-325, raw codes, one file per class
                    from .predictor import save_class_codes
                    save_class_codes(sub_class_codes, output_folder)
                few_shot_class_codes = self._gather_class_code(sub_class_codes)
                if cfg.MODEL.META_LEARN.USE_ALL_GTS_IN_BASE_CLASSES:
                    base_loader = self.build_episodic_learning_detection_test_support_set_base_loader(cfg, dataset_name)
                    chunks = []
                    for inputs in base_loader:
                        assert len(inputs) == 1, "inputs' batch size is not 1"      # :163
                        chunks.append(inputs[0])
                    base_class_codes = self._gather_class_code(inference_on_support_set_base(model, chunks), reduce=True)
                    class_codes = replace_class_code(few_shot_class_codes, base_class_codes, device=model.device)
                else:
                    class_codes = few_shot_class_codes
                class_codes = inference_normalization(model, class_codes)
                classes = self.thing_classes(cfg, dataset_name)
                if classes is not None:
                    assert len(class_codes) == len(classes), \
                        f"Got {len(class_codes)} class codes for prediction, but expect to be {len(classes)}."
                packed = format_class_codes_shared(class_codes, device=model.device)
                query_loader = self.build_episodic_learning_detection_test_query_loader(cfg, dataset_name)
                evaluator = self.get_evaluator(cfg, dataset_name, output_folder=output_folder)
                evaluator.reset()
                for inputs in query_loader:                                          # inference_on_dataset_with_class_codes :367-470
                    outputs = inference_with_class_codes(model, list(inputs), packed, batch_size=max(len(inputs), 1))
                    evaluator.process(inputs, outputs)
                res = evaluator.evaluate()
                if not _is_main_process():
                    continue                                                           # results live on the main process only
                res = {} if res is None else res
                results[f"seed{seed}"][dataset_name] = res
                # running mean of the bbox metrics over the repetitions, AP(_r/c/f)_avg / _std at the end (:602-634)
                if "bbox" in res:
                    from copy import deepcopy
                    if seed == 0:
                        results[model_tag][dataset_name] = deepcopy(res)
                    else:
                        acc = results[model_tag][dataset_name]["bbox"]
                        for k in acc:
                            acc[k] += res["bbox"][k]
                            if seed == num_repeat_test - 1:
                                acc[k] /= num_repeat_test
                    if is_final and seed == num_repeat_test - 1:
                        for k in ("AP", "APr", "APc", "APf"):
                            vals = [results[f"seed{s}"][dataset_name]["bbox"][k] for s in range(num_repeat_test)
                                    if k in results[f"seed{s}"][dataset_name].get("bbox", {})]
                            if vals:
                                arr = np.array(vals, dtype=np.float64)
                                results[model_tag][dataset_name]["bbox"][f"{k}_avg"] = float(arr.mean())
                                results[model_tag][dataset_name]["bbox"][f"{k}_std"] = float(arr.std())
        return results

    def do_test(self, cfg, model, train_iter=None, support_items=None, query_items=None):
        """`do_test(cfg, model, train_iter=None)` (meta_fcos_runner.py:674-701): the meta-learning evaluation when
        `MODEL.META_LEARN.EPISODIC_LEARNING` is on.  With in-memory `support_items` / `query_items` instead of loader
        builders it runs ONE episode (`run_episode`) and returns its detections."""
        if support_items is not None or query_items is not None:
            return run_episode(model, support_items or [], query_items or [])
        if not cfg.MODEL.META_LEARN.EPISODIC_LEARNING:
            raise NotImplementedError("base-detector evaluation (non-episodic do_test) is not on the B200 path")
        return self._do_test_meta_learning(cfg, model, train_iter=train_iter)
