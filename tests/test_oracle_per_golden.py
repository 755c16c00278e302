"""Pins oracle/per_oracle.c to the live reference PrioritizedReplayBuffer (replay_buffer.py:25-90) and
ReplayBuffer (replay_buffer.py:5-22) through tests/golden/per_*.npz / uniform_*.npz."""
import os
import random

import numpy as np
import pytest

from oracle import per_oracle as PO
from tests.golden import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@pytest.mark.parametrize("name", list(cases.PER_CASES))
def test_per_sample_bit_exact(name):
    case = cases.PER_CASES[name]
    g = np.load(os.path.join(GOLD, f"per_{name}.npz"))
    p_alpha = g["p_alpha"]
    k = min(case["batch"], p_alpha.size)
    for call in range(case["calls"]):
        u = PO.uniform_draws(case["seed"] * 10 + call, k)
        b = PO.beta(1 + call)
        assert b == g["beta"][call]
        idx, w = PO.sample(p_alpha, u, b)
        np.testing.assert_array_equal(idx, g["idx"][call])                   # bit-exact indices
        assert ulp_diff(w, g["weights"][call]).max() <= 8   # float32 pow is libm-dependent, and w / w.max() carries the max's error


@pytest.mark.parametrize("name", list(cases.PER_CASES))
def test_per_push_and_update_bit_exact(name):
    case = cases.PER_CASES[name]
    g = np.load(os.path.join(GOLD, f"per_{name}.npz"))
    cap = case["capacity"]
    pri = np.zeros(cap, np.float32)
    length = pos = 0
    for _ in range(case["n"]):
        pos = PO.push(pri, length, pos)
        length = min(length + 1, cap)
    np.testing.assert_array_equal(pri, g["prio_after_push"])
    assert pos == int(g["pos_after_push"])
    m = min(case["n"], cap)
    pri[:m] = cases.per_priorities(case)[:m]
    PO.update_priorities(pri, g["upd_idx"], cases.per_td(case)[: g["upd_idx"].size])
    np.testing.assert_array_equal(pri, g["prio_after_update"])
    pos = PO.push(pri, length, pos)
    np.testing.assert_array_equal(pri, g["prio_after_push2"])
    assert pos == int(g["pos_after_push2"])


def test_pow_alpha_close_to_numpy():
    case = cases.PER_CASES["floor"]
    g = np.load(os.path.join(GOLD, "per_floor.npz"))
    mine = PO.pow_alpha(cases.per_priorities(case), 0.6)
    assert ulp_diff(mine, g["p_alpha"]).max() <= 1


def test_pairwise_sum_matches_numpy():
    rng = np.random.RandomState(3)
    for n in [1, 7, 8, 9, 127, 128, 129, 255, 1000, 4097, 65537, 100003, 1000000]:
        a = (np.abs(rng.standard_normal(n)) ** 0.6).astype(np.float32)
        assert PO.pairwise_sum(a) == a.sum(), n


@pytest.mark.parametrize("name", list(cases.UNIFORM_CASES))
def test_uniform_ring_matches_deque(name):
    """random.sample(deque, k) == random.sample(range(n), k) picks, and deque(maxlen) eviction == ring algebra."""
    case = cases.UNIFORM_CASES[name]
    g = np.load(os.path.join(GOLD, f"uniform_{name}.npz"))
    ring = PO.DequeRing(case["capacity"])
    store = np.zeros(case["capacity"], np.int64)
    for i in range(case["n"]):
        store[ring.push_slot()] = i
    assert ring.count == int(g["len"])
    random.seed(case["seed"])
    for call in range(case["calls"]):
        j = random.sample(range(ring.count), case["batch"])
        np.testing.assert_array_equal(store[ring.physical(j)], g["r_ids"][call])
