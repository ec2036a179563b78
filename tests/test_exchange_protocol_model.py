"""Exhaustive model check of the class-code exchange protocol (kernels normalize_scatter_codes_kernel /
collect_codes_kernel in csrc/kernels_codegen.cuh) over ALL interleavings of the ranks' steps.

Per rank and episode e the device executes, in stream order:
    store(e, q)   for every peer q: the rank's code row lands in q's buffer half e % HALVES     (producer kernel, stores)
    signal(e, q)  for every peer q: q's `arrived` counter += 1                                  (after fence.sys: release)
    wait(e)       blocks until the rank's own `arrived` >= (e + 1) * rows_per_episode           (collect kernel, acquire)
    read(e, c)    for every class row c: copy row c of half e % HALVES to the caller             (must still hold episode e)
Every row cell (q, half, c) is written by rank c only, in program order, so the global state is a function of the ranks'
program counters and the reachable set is small enough to enumerate.  Claims checked:
  * with two buffer halves no reachable interleaving lets a row be overwritten (by episode e + 2) before or while its
    reader copies episode e, and every row the reader copies is the one of the current episode;
  * no deadlock: every reachable state has an enabled step until all ranks are done;
  * the checker is not vacuous: ONE buffer half, or a signal that may overtake its store (no release fence), is caught."""
from collections import deque

import pytest


def _program(world, episodes, signal_before_store=False):
    ops = []
    for e in range(episodes):
        stores = [("store", e, q) for q in range(world)]
        signals = [("signal", e, q) for q in range(world)]
        ops += (signals + stores) if signal_before_store else (stores + signals)
        ops.append(("wait", e, None))
        ops += [("read", e, c) for c in range(world)]
    return ops


def _check(world, episodes, halves, signal_before_store=False):
    """Breadth-first search over program-counter vectors.  Returns (violations, deadlocks, states)."""
    prog = _program(world, episodes, signal_before_store)   # same program on every rank (one class per rank)
    n_ops = len(prog)

    def arrived(pcs, q):
        return sum(1 for r in range(world) for op in prog[:pcs[r]] if op[0] == "signal" and op[2] == q)

    def cell(pcs, q, half, c):
        """Episode tag of row c in buffer half `half` of rank q = the last store of rank c to q into that half."""
        tag = None
        for op in prog[:pcs[c]]:
            if op[0] == "store" and op[2] == q and op[1] % halves == half:
                tag = op[1]
        return tag

    start = tuple([0] * world)
    seen, todo = {start}, deque([start])
    violations, deadlocks = [], []
    while todo:
        pcs = todo.popleft()
        enabled = 0
        for r in range(world):
            if pcs[r] == n_ops:
                continue
            kind, e, x = prog[pcs[r]]
            if kind == "wait" and arrived(pcs, r) < (e + 1) * world:
                continue
            enabled += 1
            if kind == "read" and cell(pcs, r, e % halves, x) != e:
                violations.append((pcs, r, e, x, cell(pcs, r, e % halves, x)))
            nxt = tuple(p + 1 if i == r else p for i, p in enumerate(pcs))
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
        if enabled == 0 and any(p < n_ops for p in pcs):
            deadlocks.append(pcs)
    return violations, deadlocks, len(seen)


@pytest.mark.parametrize("world,episodes", [(2, 5), (3, 4)])
def test_double_buffered_exchange_is_safe_under_every_interleaving(world, episodes):
    violations, deadlocks, states = _check(world, episodes, halves=2)
    assert states > 100
    assert not violations, violations[:3]
    assert not deadlocks, deadlocks[:3]


def test_a_single_buffer_half_is_caught():
    violations, deadlocks, _ = _check(2, 3, halves=1)
    assert violations and not deadlocks
    # the failing pattern: a fast rank stores episode e + 1 into the half its peer is still copying episode e from
    assert any(got == e + 1 for _, _, e, _, got in violations)


def test_a_signal_that_overtakes_its_store_is_caught():
    violations, _, _ = _check(2, 2, halves=2, signal_before_store=True)
    assert violations      # without the release ordering (fence.sys before the atomic) a reader can copy a stale row
