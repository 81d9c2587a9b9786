"""Host logic of the row-slab decomposition (no GPU): every level is partitioned into contiguous,
non-empty, exhaustive row ranges, level-0 cuts are even, and a coarse row always lives with its
coincident fine row -- the invariants the halo exchange and the transfer kernels rely on."""
import pytest

import eq_b200 as E


@pytest.mark.parametrize("nH,world", [(2048, 2), (2048, 8), (16384, 8), (2047, 4), (201, 2), (41, 2), (1000, 3)])
def test_partition_invariants(nH, world):
    plans = [E.slab_plan(nH, world, r) for r in range(world)]
    nlev = len(plans[0])
    assert all(len(p) == nlev for p in plans) and nlev >= 1
    for l in range(nlev):
        rows = plans[0][l][2]
        assert all(p[l][2] == rows for p in plans)
        assert plans[0][l][0] == 0 and plans[-1][l][1] == rows
        for r in range(world):
            g0, g1, _ = plans[r][l]
            assert g1 > g0                                       # nobody is idle on any level
            if r + 1 < world:
                assert plans[r + 1][l][0] == g1                  # contiguous, exhaustive
        if l == 0:
            assert all(p[0][0] % 2 == 0 for p in plans)          # even cuts on the fine level
        else:
            fine_rows = plans[0][l - 1][2]
            assert rows == fine_rows // 2 + 1                    # keep even nodes plus the last one
            for r in range(world):
                g0, g1, _ = plans[r][l]
                f0, f1, _ = plans[r][l - 1]
                for I in (g0, g1 - 1):
                    fi = min(2 * I, fine_rows - 1)
                    assert f0 <= fi < f1                          # coarse row lives with its fine row
    # the hierarchy stops while every rank still has a few rows
    assert plans[0][-1][2] // world >= 2


def test_single_rank_plan_is_the_whole_grid():
    plan = E.slab_plan(2048, 1, 0)
    assert plan[0] == (0, 2048, 2048) and plan[1] == (0, 1025, 1025)


def test_bad_arguments():
    with pytest.raises(E.EqGpuError):
        E.slab_plan(2048, 2, 2)
