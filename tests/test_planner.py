"""Static HBM arena planner (b200_plan_arena) against the reference MemorySolver's own unit-test
expectations (inference-engine/tests/unit/cpu/mkldnn_memory_solver_test.cpp, transcribed in
tests/golden/memory_solver_kat.json) plus randomised no-overlap / lower-bound properties."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from dldt_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "memory_solver_kat.json")))["cases"]


def plan(boxes):
    n = len(boxes)
    b = (C.c_int64 * max(n, 1))(*[x[2] for x in boxes])
    f = (C.c_int32 * max(n, 1))(*[x[0] for x in boxes])
    l = (C.c_int32 * max(n, 1))(*[x[1] for x in boxes])
    off = (C.c_int64 * max(n, 1))()
    tot = C.c_int64()
    _capi.check(_capi.lib.b200_plan_arena(n, b, f, l, off, C.byref(tot)))
    return list(off)[:n], tot.value


def assert_no_overlap(boxes, off):
    end = max([max(b[0], b[1]) for b in boxes] + [0]) + 1
    for i in range(len(boxes)):
        for j in range(i + 1, len(boxes)):
            a, b = boxes[i], boxes[j]
            al, bl = (a[1] if a[1] >= 0 else end), (b[1] if b[1] >= 0 else end)
            disjoint_time = al < b[0] or bl < a[0]
            disjoint_mem = off[i] + a[2] <= off[j] or off[j] + b[2] <= off[i]
            assert disjoint_time or disjoint_mem, (i, j, a, b, off[i], off[j])


def max_depth(boxes):
    end = max([max(b[0], b[1]) for b in boxes] + [0]) + 1
    depth = 0
    for t in range(end + 1):
        depth = max(depth, sum(b[2] for b in boxes if b[0] <= t <= (b[1] if b[1] >= 0 else end)))
    return depth


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_memory_solver_known_answers(case):
    off, total = plan(case["boxes"])
    assert_no_overlap(case["boxes"], off)
    if case["total"] is not None:
        assert total == case["total"]
    if "pair_sums" in case:   # GetOffset: the answer is [0,2,0,2] or [2,0,2,0]
        for k in range(len(off) - 1):
            assert off[k] + off[k + 1] == case["pair_sums"]


def test_optimal_alexnet_linear_topology():
    # mkldnn_memory_solver_test.cpp:190-226: for a linear chain the arena equals the maximum live depth
    shapes = [(3, 227, 227), (96, 55, 55), (96, 55, 55), (96, 55, 55), (96, 27, 27), (256, 27, 27), (256, 27, 27), (256, 27, 27),
              (256, 13, 13), (384, 13, 13), (384, 13, 13), (384, 13, 13), (384, 13, 13), (256, 13, 13), (256, 13, 13), (256, 6, 6),
              (1, 1, 4069), (1, 1, 4069), (1, 1, 4069), (1, 1, 4069), (1, 1, 1000), (1, 1, 1000)]
    boxes = [[k, k + 1, s[0] * s[1] * s[2]] for k, s in enumerate(shapes)]
    off, total = plan(boxes)
    assert_no_overlap(boxes, off)
    assert total == max_depth(boxes)


def test_empty_and_degenerate():
    assert plan([]) == ([], 0)
    off, total = plan([[0, 0, 0]])
    assert total == 0
    off, total = plan([[3, -1, 6]])
    assert off == [0] and total == 6


@pytest.mark.parametrize("seed", range(20))
def test_random_boxes_never_overlap(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 60))
    boxes = []
    for _ in range(n):
        s = int(rng.integers(0, 40))
        e = int(s + rng.integers(0, 12)) if rng.random() > 0.1 else -1
        boxes.append([s, e, int(rng.integers(1, 1 << 20))])
    off, total = plan(boxes)
    assert_no_overlap(boxes, off)
    assert total >= max_depth(boxes)
    assert total == max(o + b[2] for o, b in zip(off, boxes))


def test_bad_arguments():
    tot = C.c_int64()
    assert _capi.lib.b200_plan_arena(1, None, None, None, None, C.byref(tot)) == _capi.ERR_INVALID
    assert b"null" in _capi.lib.b200_last_error()
