"""One pedigree over several ranks inside the library (psim_simulate_sharded, include/psim.h): every rank simulates its
share of a generation and receives what it needs of the others' (the reference's dependency: an individual needs its
parents' haplostructs, Main.java:1504-1520). Philox streams are keyed by individual, so the sharded population must be
the unsharded one - and the oracle's - bit for bit, whatever the partition and the number of ranks.

* local transport: the ranks are contexts of this process on one GPU, every one driven by its own thread (the same SPMD
  code path as NCCL, exchange by device copies) - runs on a one-GPU box;
* NCCL transport: one process per GPU (needs two GPUs; `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

import _cases
import pedigreesim_b200 as ps

pytestmark = pytest.mark.gpu

POS_RTOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(e, ploidy, lengths, centro, p0, p1):
    e.set_chromosomes(lengths, centro, [0.3] * len(lengths) if ploidy > 2 else None, [0.5] * len(lengths) if ploidy > 2 else None)
    e.set_pedigree(p0, p1)


def _run_local(world, ploidy, p0, p1, lengths, centro, partition, seed, passes=1, gather=None):
    group = ps.ShardGroup(world)
    ctxs = []
    for r in range(world):
        g = ps.PsimContext(ploidy, seed=seed)
        _setup(g, ploidy, lengths, centro, p0, p1)
        g.shard_init_local(r, group)
        ctxs.append(g)
    errors = []

    def work(g):
        try:
            for k in range(passes):
                if k:
                    g.reset(seed, 0)
                g.simulate_sharded(partition=partition)
            if gather is not None:
                g.shard_gather(gather)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(g,)) for g in ctxs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return ctxs, group


def _unsharded(ploidy, p0, p1, lengths, centro, seed):
    out = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, ploidy, seed=seed)
        _setup(e, ploidy, lengths, centro, p0, p1)
        e.simulate()
        out.append(e)
    return out


def _assert_same(ref, g, first=0, count=None):
    so, fo, po = ref.get_haplostructs(first, count)
    sg, fg, pg = g.get_haplostructs(first, count)
    assert np.array_equal(so, sg) and np.array_equal(fo, fg)
    assert np.allclose(po, pg, rtol=POS_RTOL, atol=1e-15)


PEDIGREES = {
    "ril": lambda: _cases.ril_pedigree(37, 5),
    "f2": lambda: _cases.std_pedigree("F2", 90),
    "random": lambda: _cases.random_pedigree(np.random.default_rng(5), 5, 160),
}


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", sorted(PEDIGREES))
@pytest.mark.parametrize("ploidy", [2, 4])
def test_block_partition_leaves_the_whole_population_on_every_rank(ploidy, name, world):
    p0, p1 = PEDIGREES[name]()
    lengths, centro = [1.0, 1.6, 0.7], [0.5, 0.2, 0.7]
    o, g1 = _unsharded(ploidy, p0, p1, lengths, centro, seed=31)
    ctxs, group = _run_local(world, ploidy, p0, p1, lengths, centro, ps.PARTITION_BLOCK, seed=31, passes=2)
    for g in ctxs:
        _assert_same(o, g)
        sg, fg, pg = g.get_haplostructs()
        s1, f1, q1 = g1.get_haplostructs()
        assert np.array_equal(sg, s1) and np.array_equal(fg, f1) and np.array_equal(pg, q1)   # bit-identical to one GPU
        assert (g.shard_owner() == -1).all()
        st = g.shard_stats()
        assert st["collectives"] > 0 and st["segments_received"] > 0 and st["bytes_sent"] > 0
    total = sum(g.shard_stats()["segments_received"] for g in ctxs)
    n_sim = int(s1[-1]) - int(s1[(p0 < 0).sum() * len(lengths) * ploidy])   # segments of the simulated individuals
    assert total == (world - 1) * n_sim                                     # everybody received everybody else's share, once
    # the unsharded call on a context that has just run sharded (its device lists were this rank's share) and back
    g = ctxs[0]
    g.reset(31, 0)
    g.simulate()
    for a, b in zip(g.get_haplostructs(), g1.get_haplostructs()):
        assert np.array_equal(a, b)
    for g in ctxs:
        g.close()
    group.close()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", sorted(PEDIGREES))
def test_lineage_partition_keeps_lines_at_home(name, world):
    ploidy = 4
    p0, p1 = PEDIGREES[name]()
    lengths, centro = [1.2, 0.8], [0.6, 0.1]
    o, g1 = _unsharded(ploidy, p0, p1, lengths, centro, seed=8)
    ctxs, group = _run_local(world, ploidy, p0, p1, lengths, centro, ps.PARTITION_LINEAGE, seed=8)
    level, owner_plan, share = ps.shard_plan(p0, p1, world, ps.PARTITION_LINEAGE)
    for r, g in enumerate(ctxs):
        owner = g.shard_owner()
        assert np.array_equal(owner, np.where(share | (level == 0), -1, owner_plan))
        for i in np.nonzero((owner == r) | (owner == -1))[0]:
            _assert_same(o, g, int(i), 1)
        for i in np.nonzero((owner >= 0) & (owner != r))[0][:5]:                # not here: never sent
            with pytest.raises(ps.PsimError) as e:
                g.get_haplostructs(int(i), 1)
            assert e.value.code == ps.PSIM_ESTATE
    if name == "ril":   # a selfing chain exchanges one individual (the F1 plant), SURVEY 8e
        assert all(g.shard_stats()["segments_received"] <= 2 * len(lengths) * ploidy * 4 for g in ctxs)
    for g in ctxs:
        g.close()
    group.close()


def test_gather_brings_the_last_generation_together():
    ploidy, world = 2, 3
    p0, p1 = _cases.ril_pedigree(29, 6)
    lengths, centro = [1.5, 1.0], [0.7, 0.5]
    o, _ = _unsharded(ploidy, p0, p1, lengths, centro, seed=77)
    last = np.arange(len(p0) - 29, len(p0), dtype=np.int32)
    ctxs, group = _run_local(world, ploidy, p0, p1, lengths, centro, ps.PARTITION_LINEAGE, seed=77, gather=last)
    off, pos = _cases.uniform_map(lengths, [300, 211])
    o.set_map(off, pos)
    import _oracle
    want = o.emit(_oracle.EMIT_FOUNDER)
    for g in ctxs:
        _assert_same(o, g, int(last[0]), len(last))
        assert (g.shard_owner()[last] == -1).all()
        g.set_map(off, pos)
        got = g.emit(founder=True, ind_first=int(last[0]), ind_count=len(last))["founder"]
        assert np.array_equal(got, want[:, int(last[0]) * ploidy:])
        g.close()
    group.close()


def test_sharded_calls_need_a_communicator_and_one_replicate():
    g = ps.PsimContext(2)
    g.set_chromosomes([1.0], [0.5])
    g.set_pedigree(*_cases.std_pedigree("F1", 5))
    with pytest.raises(ps.PsimError) as e:
        g.simulate_sharded()
    assert e.value.code == ps.PSIM_ESTATE and "psim_shard_init" in str(e.value)
    group = ps.ShardGroup(1)
    h = ps.PsimContext(2, replicate_count=2)
    h.set_chromosomes([1.0], [0.5])
    h.set_pedigree(*_cases.std_pedigree("F1", 5))
    h.shard_init_local(0, group)
    with pytest.raises(ps.PsimError) as e:
        h.simulate_sharded()
    assert e.value.code == ps.PSIM_EINVAL and "replicate_count" in str(e.value)
    # a world of one is the unsharded simulation
    k = ps.PsimContext(2, seed=5)
    k.set_chromosomes([1.0], [0.5])
    k.set_pedigree(*_cases.std_pedigree("F2", 20))
    k.shard_init_local(0, group)
    k.simulate_sharded(partition=ps.PARTITION_LINEAGE)
    u = ps.PsimContext(2, seed=5)
    u.set_chromosomes([1.0], [0.5])
    u.set_pedigree(*_cases.std_pedigree("F2", 20))
    u.simulate()
    for a, b in zip(k.get_haplostructs(), u.get_haplostructs()):
        assert np.array_equal(a, b)
    for x in (g, h, k, u):
        x.close()
    group.close()


@pytest.mark.parametrize("partition", [ps.PARTITION_BLOCK, ps.PARTITION_LINEAGE])
def test_more_ranks_than_individuals_and_other_models(partition):
    """Ranks with nothing to simulate in a generation still take part in its exchange; hexaploid natural pairing with
    interference goes through the same path."""
    ploidy, world = 6, 4
    p0, p1 = _cases.std_pedigree("F2", 3)           # P1 x P2 -> F1 -> three F2 plants: generations of 1 and 3 on 4 ranks
    lengths, centro = [1.1, 0.6], [0.5, 0.3]
    out = []
    for kind in ("oracle-philox", "gpu"):
        e = _cases.make_engine(kind, ploidy, seed=3, kosambi=True, natural_pairing=True)
        _setup(e, ploidy, lengths, centro, p0, p1)
        e.simulate()
        out.append(e)
    o, g1 = out
    group = ps.ShardGroup(world)
    ctxs = []
    for r in range(world):
        g = ps.PsimContext(ploidy, seed=3, kosambi=True, natural_pairing=True)
        _setup(g, ploidy, lengths, centro, p0, p1)
        g.shard_init_local(r, group)
        ctxs.append(g)
    errors = []

    def work(g):
        try:
            g.simulate_sharded(partition=partition)
            g.shard_gather(np.arange(len(p0), dtype=np.int32))
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(g,)) for g in ctxs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g in ctxs:
        _assert_same(o, g)
        g.close()
    group.close()


def test_sharded_extension_of_a_recorded_population():
    """Mode 3 (Main.java:317-361): founders, F1 and the first selfed generation come from recorded haplostructs, the later
    generations are simulated - sharded. The plan counts the recorded individuals as done on every rank."""
    ploidy, world = 2, 2
    lines = 11
    p0, p1 = _cases.ril_pedigree(lines, 5)
    lengths, centro = [1.3, 0.9], [0.6, 0.4]
    n_rec = 3 + lines
    first = _cases.make_engine("gpu", ploidy, seed=41)
    _setup(first, ploidy, lengths, centro, p0[:n_rec], p1[:n_rec])
    first.simulate()
    rec = first.get_haplostructs()
    ref = _cases.make_engine("gpu", ploidy, seed=99)
    _setup(ref, ploidy, lengths, centro, p0, p1)
    ref.set_haplostructs(0, *rec)
    ref.simulate(int(rec[1].max()))
    group = ps.ShardGroup(world)
    ctxs = []
    for r in range(world):
        g = ps.PsimContext(ploidy, seed=99)
        _setup(g, ploidy, lengths, centro, p0, p1)
        g.set_haplostructs(0, *rec)
        g.shard_init_local(r, group)
        ctxs.append(g)
    errors = []

    def work(g):
        try:
            g.simulate_sharded(int(rec[1].max()), ps.PARTITION_BLOCK)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(g,)) for g in ctxs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g in ctxs:
        for a, b in zip(ref.get_haplostructs(), g.get_haplostructs()):
            assert np.array_equal(a, b)
        g.close()
    group.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("partition", ["block", "lineage"])
def test_nccl_transport_on_two_gpus(partition, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "run_sharded.py"), "--lines", "3000", "--generations", "6",
           "--chromosomes", "4", "--partition", partition, "--check", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    res = json.loads(out.read_text().strip().splitlines()[-1])
    assert res["identical"] is True and res["world"] == 2 and res["transport"] == "nccl"
