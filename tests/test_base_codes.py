"""Base-class "all ground truths" code path and the on-disk class-code store, CPU side:
the oracle restatement against the golden vectors produced by the REFERENCE's own reduce_class_code /
replace_class_code (oracle/make_golden.py), and the host logic of the code store."""
import copy
import os

import pytest
import torch

from oracle import base_codes_oracle as bo
from tests.cases import load_golden


def _same_code(a, b):
    return torch.equal(a["cls_conv"], b["cls_conv"]) and torch.equal(a["cls_bias"], b["cls_bias"])


def test_oracle_accumulation_and_reduce_match_reference_golden():
    g = load_golden("base_reduce")
    per_rank = []
    for rk in g["chunks_per_rank"]:
        per_rank.append(bo.accumulate_base_codes([copy.deepcopy(c["code"]) for c in rk], [c["cid"] for c in rk],
                                                 [c["len"] for c in rk], [c["total_len"] for c in rk], [c["name"] for c in rk]))
    for mine, ref in zip(per_rank, g["per_rank"]):
        assert [c["support_set_target"] for c in mine] == [c["support_set_target"] for c in ref]
        for a, b in zip(mine, ref):
            assert _same_code(a["class_code"], b["class_code"])
            assert a["class_code"]["acc_weight"] == b["class_code"]["acc_weight"]
    reduced = bo.reduce_class_code([c for rk in per_rank for c in rk])
    assert len(reduced) == len(g["reduced"]) == 6
    for a, b in zip(reduced, g["reduced"]):
        assert bo._cid(a["support_set_target"]) == bo._cid(b["support_set_target"]) and a["class_name"] == b["class_name"]
        assert _same_code(a["class_code"], b["class_code"]) and "acc_weight" not in a["class_code"]
    replaced = bo.replace_class_code(g["few_shot"], reduced)
    for a, b in zip(replaced, g["replaced"]):
        assert _same_code(a["class_code"], b["class_code"])
    # ids 6 and 7 have no base code: untouched
    assert _same_code(replaced[6]["class_code"], g["few_shot"][6]["class_code"])


def test_rebalance_only_when_weight_is_not_one():
    g = load_golden("base_reduce")
    weights = {}
    for rk in g["per_rank"]:
        for c in rk:
            weights[c["support_set_target"]] = weights.get(c["support_set_target"], 0) + c["class_code"]["acc_weight"]
    assert abs(weights[4] - 1.0) > 1e-6          # class 4 lost its last chunk -> divided by acc_weight
    assert all(abs(weights[c] - 1.0) <= 1e-6 for c in (0, 1, 2, 3, 5))


def test_empty_inputs():
    assert bo.reduce_class_code([]) == []
    from sylph_few_shot_detection_b200.runner import reduce_class_code
    assert reduce_class_code([]) == []


def test_reduce_refuses_to_run_without_the_device():
    from sylph_few_shot_detection_b200.runner import reduce_class_code
    g = load_golden("base_reduce")
    with pytest.raises(RuntimeError):
        reduce_class_code([c for rk in g["per_rank"] for c in rk], engine=None)


def test_replace_class_code_host_logic():
    from sylph_few_shot_detection_b200.runner import replace_class_code
    g = load_golden("base_reduce")
    out = replace_class_code(g["few_shot"], g["reduced"], "cpu")
    for a, b in zip(out, g["replaced"]):
        assert _same_code(a["class_code"], b["class_code"])


def test_class_code_store_roundtrip(tmp_path):
    from sylph_few_shot_detection_b200.predictor import load_class_code_list, resize_shortest_edge_shape, save_class_codes
    g = torch.Generator().manual_seed(0)
    items = [{"support_set": ["dropped"], "support_set_target": torch.tensor(i), "class_name": f"cls {i}",
              "class_code": {"cls_conv": torch.randn(1, 256, 1, 1, generator=g), "cls_bias": torch.randn(1, 1, 1, 1, generator=g)}}
             for i in range(3)]
    paths = save_class_codes(items, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["cls 0.pth", "cls 1.pth", "cls 2.pth"]
    save_class_codes(items[:1], str(tmp_path))                       # overwriting an existing file is allowed
    back = load_class_code_list(str(tmp_path), ["cls 2", "cls 0"])
    assert "support_set" not in back[0] and back[0]["class_name"] == "cls 2" and int(back[0]["support_set_target"]) == 2
    assert torch.equal(back[1]["class_code"]["cls_conv"], items[0]["class_code"]["cls_conv"])
    with pytest.raises(ValueError, match="is missing"):
        load_class_code_list(str(tmp_path), ["cls 0", "nope"])
    # reference file layout: a plain torch.save of the dict (meta_learn_evaluation.py:322-325)
    raw = torch.load(os.path.join(str(tmp_path), "cls 1.pth"), weights_only=False)
    assert set(raw) == {"support_set_target", "class_name", "class_code"} and set(raw["class_code"]) == {"cls_conv", "cls_bias"}
    assert resize_shortest_edge_shape(480, 640, 800, 1333) == (800, 1067)
    assert resize_shortest_edge_shape(500, 2000, 800, 1333) == (333, 1333)
    assert resize_shortest_edge_shape(800, 1333, 800, 1333) == (800, 1333)
