"""Host logic of the pinned host-buffer batch call (csrc/tvl1_solver.cu: ramp_schedule), no GPU needed: how
tvl1_solve_batch_* / tvl1_solve_sequence_* cut a batch into lock-step chunks (include/tvl1_b200.h: tvl1_plan_chunks)."""
import pytest

import optical_flow_1_b200 as pkg
from optical_flow_1_b200 import tvl1


@pytest.mark.parametrize("npairs", [1, 2, 3, 7, 16, 17, 24, 40, 100, 129, 255, 256, 257, 1000, 4099])
@pytest.mark.parametrize("max_batch", [1, 2, 3, 4, 6, 16, 32, 48, 64, 256])
def test_plan_covers_the_batch_with_few_sizes(npairs, max_batch):
    c = tvl1.plan_chunks(npairs, max_batch)
    assert sum(c) == npairs and min(c) >= 1 and max(c) <= max_batch
    # a lane keeps one workspace + solve graph per chunk size: the current one and four alternates
    assert len(set(c)) <= 5, c
    if npairs <= max_batch:
        assert c == [npairs]


def test_plan_ramps_at_both_ends():
    c = tvl1.plan_chunks(256, 64)
    assert c == [8, 16, 32, 64, 64, 32, 16, 16, 8]
    # no ramp step below 8 pairs (very small lock-step chunks: DESIGN 3.6)
    assert tvl1.plan_chunks(256, 16) == [8] + [16] * 15 + [8]
    assert tvl1.plan_chunks(24, 4) == [4] * 6
    c = tvl1.plan_chunks(1000, 64)
    assert c[:4] == [8, 16, 32, 64] and c[-1] == 8 and c.count(64) == 13
    # the tail never grows again
    k = max(i for i, b in enumerate(c) if b == 64)
    assert all(c[i] >= c[i + 1] for i in range(k, len(c) - 1)), c
    # too small for full ramps: shorter ramps, never a chunk above max_batch
    assert tvl1.plan_chunks(129, 64) == [32, 64, 32, 1]
    assert tvl1.plan_chunks(17, 16) == [16, 1]


def test_plan_rejects_nonsense():
    assert tvl1.plan_chunks(0, 8) == [] and tvl1.plan_chunks(8, 0) == []
