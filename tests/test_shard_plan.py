"""psim_shard_plan (the partition psim_simulate_sharded uses; pure host code of libpsim, no device) against the numpy
statement of the same rules in pedigreesim_b200/sharded.py, and its invariants: every individual that a child needs is
either on the child's rank or shared after its own generation (the reference's dependency, Main.java:1504-1520)."""
import numpy as np
import pytest

import _cases
import pedigreesim_b200 as ps
from pedigreesim_b200 import sharded


def _pedigrees():
    rng = np.random.default_rng(11)
    out = {
        "f1": _cases.std_pedigree("F1", 41),
        "f2": _cases.std_pedigree("F2", 30),
        "bc": _cases.std_pedigree("BC", 17),
        "ril": _cases.ril_pedigree(23, 6),
        "random": _cases.random_pedigree(rng, 6, 300),
        "deep": _cases.random_pedigree(rng, 3, 120),
    }
    # two unrelated families side by side + a late cross between them
    a0, a1 = _cases.ril_pedigree(5, 4)
    n = len(a0)
    b0 = np.where(a0 >= 0, a0 + n, -1)
    b1 = np.where(a1 >= 0, a1 + n, -1)
    p0 = np.concatenate([a0, b0, [n - 1]]).astype(np.int32)
    p1 = np.concatenate([a1, b1, [2 * n - 1]]).astype(np.int32)
    out["two_families"] = (p0, p1)
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("name", sorted(_pedigrees()))
def test_plan_matches_the_numpy_statement(name, world):
    p0, p1 = _pedigrees()[name]
    levels = sharded.levelize(p0, p1)
    level, owner, share = ps.shard_plan(p0, p1, world, ps.PARTITION_LINEAGE)
    want_level = np.zeros(len(p0), np.int32)
    for lv, inds in enumerate(levels):
        want_level[inds] = lv + 1
    assert np.array_equal(level, want_level)
    want_owner, want_share = sharded.lineage_partition(levels, p0, p1, world)
    assert np.array_equal(owner, want_owner)
    assert np.array_equal(share, want_share)
    level_b, owner_b, share_b = ps.shard_plan(p0, p1, world, ps.PARTITION_BLOCK)
    assert np.array_equal(level_b, want_level)
    for inds in levels:
        for r, blk in enumerate(sharded.block_partition(inds, world)):
            assert (owner_b[blk] == r).all()
        assert share_b[inds].all()
    assert (owner_b[want_level == 0] == -1).all() and not share_b[want_level == 0].any()


@pytest.mark.parametrize("world", [2, 4, 5])
@pytest.mark.parametrize("name", sorted(_pedigrees()))
def test_every_parent_is_where_its_child_needs_it(name, world):
    p0, p1 = _pedigrees()[name]
    level, owner, share = ps.shard_plan(p0, p1, world, ps.PARTITION_LINEAGE)
    for i in np.nonzero(level > 0)[0]:
        assert 0 <= owner[i] < world
        for par in (p0[i], p1[i]):
            assert level[par] < level[i]
            assert level[par] == 0 or share[par] or owner[par] == owner[i], (i, par)
    # a selfing chain or an F1 shares next to nothing
    if name in ("f1", "ril"):
        assert share.sum() <= 1
    # the load is balanced within a generation's rounding for pedigrees whose lines can be dealt out freely
    if name in ("f1", "ril", "f2"):
        cnt = np.bincount(owner[level > 0], minlength=world)
        assert cnt.max() - cnt.min() <= 2 * level.max()


def test_plan_with_individuals_that_are_already_simulated():
    p0, p1 = _cases.ril_pedigree(8, 5)
    has = np.zeros(len(p0), bool)
    has[:3 + 8] = True          # founders, F1 and the first selfed generation exist already (mode-3 extension)
    level, owner, share = ps.shard_plan(p0, p1, 2, ps.PARTITION_LINEAGE, has_haplostruct=has)
    assert (level[has] == 0).all() and (owner[has] == -1).all()
    levels = sharded.levelize(p0, p1, has)
    want_owner, want_share = sharded.lineage_partition(levels, p0, p1, 2)
    assert np.array_equal(owner, want_owner) and np.array_equal(share, want_share)


def test_plan_rejects_what_the_library_rejects():
    with pytest.raises(ps.PsimError):
        ps.shard_plan([-1, 2, 0], [-1, 2, 0], 2, ps.PARTITION_BLOCK)        # child before its parent
    with pytest.raises(ps.PsimError):
        ps.shard_plan([-1, 0], [-1, -1], 2, ps.PARTITION_BLOCK)             # one known and one unknown parent
    with pytest.raises(ps.PsimError):
        ps.shard_plan([-1, -1, 0], [-1, -1, 1], 17, ps.PARTITION_BLOCK)     # more ranks than PSIM_SHARD_MAX_WORLD
    with pytest.raises(ps.PsimError):
        ps.shard_plan([-1, -1, 0], [-1, -1, 1], 2, 7)                       # unknown partition
