"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the base-class "all ground truths" code path of the reference:
  * accumulate_base_codes ... the per-chunk weighted accumulation inside inference_on_support_set_dataset_base,
                              sylph/evaluation/meta_learn_evaluation.py:190-203, restated line for line; pinned against the
                              reference's real loop (imported unmodified on evaluator stand-ins) by
                              tests/test_reference_loops_dropin.py: bit-equal codes and acc_weight;
  * reduce_class_code ....... sylph/modeling/code_generator/utils.py:397-427 (+ convert_list_to_dict :340-374);
  * replace_class_code ...... sylph/modeling/code_generator/utils.py:376-394.
Pinned: `oracle/make_golden.py` runs the reference's own reduce_class_code / replace_class_code (imported unmodified
through oracle/reference_loader.py) on seeded inputs, checks this restatement against them bit for bit and stores the
vectors in tests/golden/base_reduce.pt.
"""
from __future__ import annotations

import functools
from collections import OrderedDict
from typing import Dict, List, Sequence

import torch


def _cid(t) -> int:
    return int(t.item()) if torch.is_tensor(t) else int(t)


def accumulate_base_codes(chunk_codes: Sequence[Dict[str, torch.Tensor]], cids: Sequence[int], lens: Sequence[int],
                          total_lens: Sequence[int], names: Sequence[str]) -> List[Dict]:
    """meta_learn_evaluation.py:181-224: weight = float(len) / total_len; first chunk of a class assigns
    code * weight, later chunks add code * weight; acc_weight accumulates as a Python float."""
    cid_to_class_code: "OrderedDict[int, Dict]" = OrderedDict()
    cid_to_class_name: Dict[int, str] = {}
    for code, cid, ln, tot, name in zip(chunk_codes, cids, lens, total_lens, names):
        cid_to_class_name[cid] = name
        weight = float(ln) / tot
        if cid in cid_to_class_code:
            cid_to_class_code[cid]["cls_conv"] += code["cls_conv"] * weight
            cid_to_class_code[cid]["cls_bias"] += code["cls_bias"] * weight
            cid_to_class_code[cid]["acc_weight"] += weight
        else:
            cid_to_class_code[cid] = {"cls_conv": code["cls_conv"] * weight, "cls_bias": code["cls_bias"] * weight,
                                      "acc_weight": weight}
    return [{"support_set_target": cid, "class_name": cid_to_class_name[cid], "class_code": cid_to_class_code[cid]}
            for cid in cid_to_class_code]


def reduce_class_code(out_codes: List[Dict]) -> List[Dict]:
    """utils.py:397-427 (without the logging and without the final log line that requires cls_weight_norm)."""
    if len(out_codes) == 0:
        return out_codes
    all_keys = out_codes[0]["class_code"].keys()
    lst: "OrderedDict[int, List[Dict]]" = OrderedDict()
    other: Dict[int, Dict] = {}
    for c in out_codes:
        cid = _cid(c["support_set_target"])
        lst.setdefault(cid, []).append(c["class_code"])
        if cid not in other:
            other[cid] = {k: v for k, v in c.items() if k != "class_code"}
    results = []
    for cid, code_lst in lst.items():
        result = other[cid]
        result["class_code"] = {key: functools.reduce(lambda x, y: x + y[key], code_lst, 0) for key in all_keys}
        acc_weight = result["class_code"]["acc_weight"]
        if abs(1.0 - acc_weight) > 1e-6:
            result["class_code"]["cls_conv"] = result["class_code"]["cls_conv"] / acc_weight
            result["class_code"]["cls_bias"] = result["class_code"]["cls_bias"] / acc_weight
            result["class_code"]["acc_weight"] = 1.0
        del result["class_code"]["acc_weight"]
        results.append(result)
    return results


def replace_class_code(support_set_class_code: List[Dict], target_class_codes: List[Dict]) -> List[Dict]:
    """utils.py:376-394: overlapping class ids take the FIRST target code of that id."""
    target: Dict[int, Dict] = {}
    for c in target_class_codes:
        target.setdefault(_cid(c["support_set_target"]), c["class_code"])
    out = []
    for c in support_set_class_code:
        r = dict(c)
        cid = _cid(c["support_set_target"])
        r["class_code"] = dict(target[cid]) if cid in target else dict(c["class_code"])
        out.append(r)
    return out
