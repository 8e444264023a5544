"""Host logic of the multi-GPU partitioner: the order-preserving repartition plan
(pgn_plan_rebalance, csrc/comm.cc).  Pure host code: runs without a GPU."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import gpuintegration_b200 as pg


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=10_000), min_size=1, max_size=8))
def test_plan_is_balanced_complete_and_order_preserving(counts):
    new_counts, transfers = pg.plan_rebalance(counts)
    w, total = len(counts), sum(counts)
    assert sum(new_counts) == total
    assert max(new_counts) - min(new_counts) <= 1
    assert all(new_counts[i] >= new_counts[i + 1] for i in range(w - 1))
    # apply the plan to labelled items: rank-major concatenation must be unchanged
    src = [np.arange(sum(counts[:r]), sum(counts[:r + 1])) for r in range(w)]
    dst = [np.full(new_counts[r], -1) for r in range(w)]
    for s, d, so, do, cnt in transfers:
        assert cnt > 0 and so + cnt <= counts[s] and do + cnt <= new_counts[d]
        assert np.all(dst[d][do:do + cnt] == -1), "overlapping destination ranges"
        dst[d][do:do + cnt] = src[s][so:so + cnt]
    got = np.concatenate(dst) if total else np.array([], dtype=int)
    assert np.array_equal(got, np.arange(total))
    # every source item moves exactly once
    moved = sum(cnt for *_x, cnt in transfers)
    assert moved == total
    # a contiguous repartition never needs more than 2w - 1 runs
    assert len(transfers) <= 2 * w


def test_plan_examples():
    assert pg.plan_rebalance([10, 0, 5, 1])[0] == [4, 4, 4, 4]
    nc, tr = pg.plan_rebalance([4, 4])
    assert nc == [4, 4] and tr == [(0, 0, 0, 0, 4), (1, 1, 0, 0, 4)]
    nc, tr = pg.plan_rebalance([0, 0, 7])
    assert nc == [3, 2, 2]
    assert tr == [(2, 0, 0, 0, 3), (2, 1, 3, 0, 2), (2, 2, 5, 0, 2)]
    assert pg.plan_rebalance([0, 0]) == ([0, 0], [])
