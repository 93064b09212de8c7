"""Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo process groups.

* one large pedigree: generation blocks + variable-length all-gather of the new HaploStructs
  (pedigreesim_b200/sharded.py) against a single-process run, with a deterministic stand-in for
  the meiosis kernels (a child's HaploStructs are a pure function of its parents', so a wrong or
  missing exchange changes the result);
* replicate pedigrees: the replicate ids of bench.py's weak-scaling mode partition the work with
  no overlap and no gap.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _cases
from pedigreesim_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeShard:
    """CPU stand-in with the shard interface; haplostructs are numpy CSR per individual."""

    def __init__(self, p0, p1, C=2, P=4):
        self.p0, self.p1, self.C, self.P = p0, p1, C, P
        self.hs = {}  # individual -> list over (c, h) of (founder[], pos[])
        self.next_allele = 0

    def simulate_subset(self, individuals, maxfounderallele=-1):
        for i in range(len(self.p0)):       # founders, in pedigree order like Main.java:1510-1513
            if self.p0[i] < 0 and i not in self.hs:
                self.hs[i] = [([self.next_allele + h], [0.0]) for _c in range(self.C) for h in range(self.P)]
                self.next_allele += self.P
        for i in individuals:
            i = int(i)
            out = []
            for c in range(self.C):
                for h in range(self.P):
                    par = self.p0[i] if h < self.P // 2 else self.p1[i]
                    f, x = self.hs[int(par)][c * self.P + (i + h) % self.P]    # KeyError if the exchange failed
                    out.append((list(f) + [f[-1] + i % 3], list(x) + [x[-1] + 0.01 * (1 + (i + c + h) % 7)]))
            self.hs[i] = out

    def export_range(self, first, count):
        off, fo, po = [0], [], []
        for i in range(first, first + count):
            for f, x in self.hs[i]:
                fo += f
                po += x
                off.append(len(fo))
        return (torch.tensor(off, dtype=torch.int64), torch.tensor(fo, dtype=torch.int32), torch.tensor(po, dtype=torch.float64))

    def import_range(self, first, count, off, fo, po):
        off, fo, po = off.tolist(), fo.tolist(), po.tolist()
        k = 0
        for i in range(first, first + count):
            hs = []
            for _ in range(self.C * self.P):
                hs.append((fo[off[k]:off[k + 1]], po[off[k]:off[k + 1]]))
                k += 1
            self.hs[i] = hs


def _pedigrees():
    rng = np.random.default_rng(1)
    return {
        "f1": _cases.std_pedigree("F1", 37),
        "ril": _cases.ril_pedigree(9, 5),
        "random": _cases.random_pedigree(rng, 5, 60),
    }


def _worker(rank, world, port, name, q, partition="block"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p0, p1 = _pedigrees()[name]
    shard = FakeShard(p0, p1)
    drv = sharded.ShardedPedigree(shard, p0, p1, partition=partition)
    drv.simulate()
    if partition == "block":
        q.put((rank, {i: shard.hs[i] for i in range(len(p0))}, drv.exchanged_segments))
    else:
        before = dict(shard.hs)
        collectives = drv.collectives
        owned = drv.owned().tolist()
        last = sharded.levelize(p0, p1)[-1]
        drv.gather(last)                       # e.g. for a joint emission of the last generation
        q.put((rank, before, collectives, owned, {int(i): shard.hs[int(i)] for i in last}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["f1", "ril", "random"])
def test_sharded_pedigree_matches_single_process(name):
    p0, p1 = _pedigrees()[name]
    ref = FakeShard(p0, p1)
    ref.simulate_subset(np.zeros(0, np.int32))
    for lvl in sharded.levelize(p0, p1):
        ref.simulate_subset(lvl)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, hs, exchanged in results:
        assert hs == ref.hs, "rank %d diverged" % rank
        assert exchanged > 0


def test_levelize_and_partition():
    p0, p1 = _cases.ril_pedigree(4, 3)
    lv = sharded.levelize(p0, p1)
    assert [l.tolist() for l in lv] == [[2], [3, 4, 5, 6], [7, 8, 9, 10]]
    parts = sharded.block_partition(np.arange(10), 4)
    assert [p.tolist() for p in parts] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert sharded.contiguous_runs([3, 4, 5, 9, 10, 12]) == [(3, 3), (9, 2), (12, 1)]
    has = np.zeros(len(p0), bool)
    has[:7] = True   # extend mode: generations up to F2 preset
    assert [l.tolist() for l in sharded.levelize(p0, p1, has)] == [[7, 8, 9, 10]]
    with pytest.raises(ValueError):
        sharded.levelize([-1, 2, 0], [-1, 2, 0])


@pytest.mark.parametrize("per_step", [1, 8])
def test_replicate_ids_partition_without_overlap(per_step):
    # bench.py: in step k, rank r of `world` simulates the `per_step` replicate ids from (r + world * (k + 1)) * per_step on
    # (one context with replicate_count = per_step)
    world, steps = 4, 5
    seen = [(r + world * (k + 1)) * per_step + j for r in range(world) for k in range(steps) for j in range(per_step)]
    assert len(set(seen)) == world * steps * per_step
    assert sorted(seen) == list(range(world * per_step, world * (steps + 1) * per_step))


@pytest.mark.parametrize("name", ["f1", "ril", "random"])
def test_lineage_partition_exchanges_only_what_another_rank_needs(name):
    """partition="lineage": a rank keeps its lines; F1-from-founders needs no collective at all, a
    selfing chain one (its single F1 plant); every rank's individuals equal the single-process run,
    the owned sets partition the non-founders, and gather() completes a generation everywhere."""
    p0, p1 = _pedigrees()[name]
    ref = FakeShard(p0, p1)
    ref.simulate_subset(np.zeros(0, np.int32))
    levels = sharded.levelize(p0, p1)
    for lvl in levels:
        ref.simulate_subset(lvl)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q, "lineage")) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned_all = []
    for rank, hs, collectives, owned, last_gen in results:
        for i, v in hs.items():
            assert v == ref.hs[i], "rank %d individual %d diverged" % (rank, i)
        assert set(owned) <= set(hs)
        owned_all += owned
        assert last_gen == {int(i): ref.hs[int(i)] for i in levels[-1]}
        if name == "f1":
            assert collectives == 0
        if name == "ril":
            assert collectives == 1 and len(hs) < len(p0)     # only the F1 plant was shared
    assert sorted(owned_all) == [i for i in range(len(p0)) if p0[i] >= 0]


def test_lineage_partition_invariants():
    rng = np.random.default_rng(3)
    for p0, p1 in (_cases.ril_pedigree(1000, 6), _cases.std_pedigree("F2", 501), _cases.random_pedigree(rng, 12, 800)):
        levels = sharded.levelize(p0, p1)
        for world in (2, 3, 8):
            owner, share = sharded.lineage_partition(levels, p0, p1, world)
            assert (owner[p0 < 0] == -1).all() and (owner[p0 >= 0] >= 0).all() and owner.max() < world
            for inds in levels:
                for par in (p0[inds], p1[inds]):
                    assert ((owner[par] == -1) | (owner[par] == owner[inds]) | share[par]).all()
    p0, p1 = _cases.ril_pedigree(1000, 6)
    owner, share = sharded.lineage_partition(sharded.levelize(p0, p1), p0, p1, 8)
    load = np.bincount(owner[owner >= 0], minlength=8)
    assert load.max() - load.min() <= 1 + 1 and share.sum() == 1   # balanced, only the F1 plant is shared
