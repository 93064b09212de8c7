"""One large pedigree over several GPUs.

One process per GPU. Every rank holds the whole pedigree in its own libpsim context but simulates
only its block of each generation; after a generation the new HaploStructs are exchanged with one
variable-length all-gather (sizes, then the packed CSR) so that any rank can read any parent in the
next generation. Replicate pedigrees need none of this (no data-path collective: each rank just
uses its own replicate ids).

Two partitions. "block": the individuals of each generation are cut into equal blocks and every new
HaploStruct is shared with everybody (any pedigree; every rank ends with the whole population).
"lineage": an individual lives on the rank that owns its parents whenever that is possible, and only
individuals with a child on ANOTHER rank are shared - for an F1 from founders or a selfing chain
 nothing is exchanged at all and every rank ends with its own lines (emission needs
no exchange either: a rank emits the individuals it owns).

The exchange is written against `torch.distributed` only, so the same code runs over NCCL / NVLink
on device tensors and over gloo on CPU tensors (the world-size-2 tests).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------ host logic
def levelize(parent0, parent1, has_hs=None):
    """Generations of a parents-first pedigree: level = 1 + max(level of parents); founders and
    individuals that already have a haplostruct are level 0. Returns a list of index arrays
    (level 1, 2, ...). The reference walks the individuals strictly in order
    (Main.simulateIndividuals, Main.java:1504-1520); levels are the units that can run in parallel."""
    p0 = np.asarray(parent0).astype(np.int64)
    p1 = np.asarray(parent1).astype(np.int64)
    n = len(p0)
    idx = np.arange(n)
    todo = p0 >= 0
    if has_hs is not None:
        todo &= ~np.asarray(has_hs, bool)
    if np.any(todo & ((p0 >= idx) | (p1 >= idx) | (p1 < 0))):
        raise ValueError("pedigree is not in parents-first order")
    level = np.zeros(n, np.int32)
    t = np.nonzero(todo)[0]
    while True:   # one relaxation per generation (vectorised; parents-first order bounds the depth)
        new = 1 + np.maximum(level[p0[t]], level[p1[t]])
        if np.array_equal(new, level[t]):
            break
        level[t] = new
    return [np.nonzero(level == l)[0].astype(np.int32) for l in range(1, int(level.max()) + 1)]


def block_partition(items, world):
    """Contiguous, near-equal blocks of `items` for ranks 0..world-1."""
    items = np.asarray(items)
    base, extra = divmod(len(items), world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append(items[start:start + cnt])
        start += cnt
    return out


def lineage_partition(levels, parent0, parent1, world):
    """Owner rank of every individual for the "lineage" partition, and which individuals must be
    shared. Founders and individuals that already have a haplostruct are everywhere (owner -1).
    Within a generation: a child whose parents are both everywhere is free (the free ones are dealt
    out in contiguous blocks, lightest rank first); otherwise it goes to the owner of its first owned
    parent. A parent with many children in one generation is shared rather than followed. Returns (owner int32 [n], share bool [n]): share[i] means i has a child on another rank and
    is all-gathered right after its generation (it is everywhere from then on)."""
    p0 = np.asarray(parent0).astype(np.int64)
    p1 = np.asarray(parent1).astype(np.int64)
    n = len(p0)
    owner = np.full(n, -1, np.int32)
    share = np.zeros(n, bool)
    load = np.zeros(world, np.int64)
    for inds in levels:
        inds = np.asarray(inds, np.int64)
        # a bottleneck parent (the one F1 plant of a RIL chain, a popular sire) would pull its whole
        # progeny onto one rank: parents with many children in this generation are shared instead
        heavy = max(2.0, len(inds) / (4.0 * world))
        for par in (p0[inds], p1[inds]):
            owned = par[(owner[par] >= 0) & ~share[par]]
            if len(owned):
                ids, kids = np.unique(owned, return_counts=True)
                share[ids[kids > heavy]] = True
        a, b = owner[p0[inds]], owner[p1[inds]]
        a = np.where(share[p0[inds]], -1, a)          # a shared parent is everywhere
        b = np.where(share[p1[inds]], -1, b)
        own = np.where(a >= 0, a, b)                  # first owned parent, -1 if both are everywhere
        free = inds[own < 0]
        fixed = own >= 0
        owner[inds[fixed]] = own[fixed]
        load += np.bincount(own[fixed], minlength=world)
        # the other parent lives elsewhere: it has to be shared before this generation runs. It
        # belongs to an EARLIER generation whose exchange is over, so the decision is taken in a
        # first pass over the levels (see ShardedPedigree): here it only marks.
        clash = fixed & (a >= 0) & (b >= 0) & (a != b)
        share[p1[inds[clash]]] = True
        if len(free) and load.max() == load.min():    # the usual case: equal blocks
            for r, blk in enumerate(block_partition(free, world)):
                owner[blk] = r
                load[r] += len(blk)
        elif len(free):
            target = (load.sum() + len(free)) / world
            start = 0
            for r in np.argsort(load, kind="stable"):
                cnt = int(min(len(free) - start, max(0, round(target - load[r]))))
                owner[free[start:start + cnt]] = r
                load[r] += cnt
                start += cnt
            if start < len(free):                     # rounding leftovers
                r = int(np.argmin(load))
                owner[free[start:]] = r
                load[r] += len(free) - start
    return owner, share


def contiguous_runs(indices):
    """Split a sorted index array into (first, count) runs of consecutive indices."""
    idx = np.asarray(indices)
    if len(idx) == 0:
        return []
    cuts = np.nonzero(np.diff(idx) != 1)[0] + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(idx)]])
    return [(int(idx[a]), int(b - a)) for a, b in zip(starts, ends)]


# ------------------------------------------------------------------------------------ exchange
def allgather_csr(seg_off, founder, pos, group=None):
    """Variable-length all-gather of a packed CSR block (seg_off int64 [h+1] starting at 0,
    founder int32 [s], pos float64 [s]); tensors may live on the GPU (NCCL) or the CPU (gloo).
    Returns a list with one (seg_off, founder, pos) triple per rank."""
    world = dist.get_world_size(group)
    dev = seg_off.device
    sizes = torch.tensor([seg_off.numel(), founder.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [tuple(int(v) for v in s.tolist()) for s in all_sizes]
    max_h = max(s[0] for s in all_sizes)
    max_s = max(max(s[1] for s in all_sizes), 1)

    def padded(t, n):
        out = torch.zeros(n, dtype=t.dtype, device=dev)
        out[:t.numel()] = t
        return out

    g_off = [torch.empty(max_h, dtype=torch.int64, device=dev) for _ in range(world)]
    g_f = [torch.empty(max_s, dtype=torch.int32, device=dev) for _ in range(world)]
    g_p = [torch.empty(max_s, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(g_off, padded(seg_off, max_h), group=group)
    dist.all_gather(g_f, padded(founder, max_s), group=group)
    dist.all_gather(g_p, padded(pos, max_s), group=group)
    return [(g_off[r][:all_sizes[r][0]], g_f[r][:all_sizes[r][1]], g_p[r][:all_sizes[r][1]]) for r in range(world)]


class _DevArray:
    """Minimal __cuda_array_interface__ wrapper around a raw device pointer owned by libpsim."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PsimShard:
    """Adapter: range export / import of a PsimContext as torch tensors on its device."""

    def __init__(self, ctx):
        self.ctx = ctx

    def simulate_subset(self, individuals, maxfounderallele=-1):
        self.ctx.simulate_subset(individuals, maxfounderallele)

    def export_range(self, first, count):
        d = self.ctx.export_device(first, count)
        nh, ns = int(d.n_homologs), int(d.n_segments)
        off = torch.as_tensor(_DevArray(d.seg_off, nh + 1, "<i8"), device="cuda").clone()
        fo = torch.as_tensor(_DevArray(d.founder, max(ns, 1), "<i4"), device="cuda")[:ns].clone()
        po = torch.as_tensor(_DevArray(d.pos, max(ns, 1), "<f8"), device="cuda")[:ns].clone()
        return off, fo, po

    def import_range(self, first, count, off, fo, po):
        from ._lib import DeviceHS
        off, fo, po = off.contiguous(), fo.contiguous(), po.contiguous()
        # libpsim copies on its own stream: the collective that filled these tensors (enqueued on
        # torch's stream) must have finished first
        torch.cuda.current_stream().synchronize()
        d = DeviceHS(off.numel() - 1, fo.numel(), C.c_void_p(off.data_ptr()), C.c_void_p(fo.data_ptr()), C.c_void_p(po.data_ptr()))
        self.ctx.import_device(first, count, d)


class ShardedPedigree:
    """Generation-by-generation driver for one pedigree sharded over the ranks of a process group.

    `shard` is any object with simulate_subset(individuals, maxfounderallele), export_range(first,
    count) -> (seg_off, founder, pos) tensors and import_range(first, count, seg_off, founder, pos).
    """

    def __init__(self, shard, parent0, parent1, group=None, partition="block"):
        if partition not in ("block", "lineage"):
            raise ValueError("partition must be 'block' or 'lineage'")
        self.shard = shard
        self.p0 = np.asarray(parent0, np.int32)
        self.p1 = np.asarray(parent1, np.int32)
        self.group = group
        self.partition = partition
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.exchanged_segments = 0
        self.collectives = 0
        self.owner = None     # lineage partition: owner rank per individual (-1: everywhere)
        self._levels = None   # generations of the whole pedigree (has_hs None): computed once, the pedigree is fixed
        self._plan = None     # per generation: (my individuals, per-rank lists of individuals to share)

    def _make_plan(self, levels):
        plan = []
        if self.partition == "block":
            for inds in levels:
                blocks = block_partition(inds, self.world)
                plan.append((blocks[self.rank], blocks))
        else:
            owner, share = lineage_partition(levels, self.p0, self.p1, self.world)
            self.owner = owner
            for inds in levels:
                mine = inds[owner[inds] == self.rank]
                shared = [inds[(owner[inds] == r) & share[inds]] for r in range(self.world)]
                plan.append((mine, shared))
        return plan

    def owned(self):
        """Individuals this rank holds a haplostruct of after simulate() that nobody sent it: its own
        share of every generation (lineage partition; with the block partition every rank ends with
        everything)."""
        if self.owner is None:
            return None
        return np.nonzero(self.owner == self.rank)[0].astype(np.int32)

    def simulate(self, maxfounderallele=-1, has_hs=None):
        if has_hs is None:
            if self._levels is None:
                self._levels = levelize(self.p0, self.p1)
                self._plan = self._make_plan(self._levels)
            plan = self._plan
        else:
            plan = self._make_plan(levelize(self.p0, self.p1, has_hs))
        # founders are initialised by every rank (deterministic, no exchange needed)
        self.shard.simulate_subset(np.zeros(0, np.int32), maxfounderallele)
        for mine, blocks in plan:
            if len(mine):
                self.shard.simulate_subset(mine, maxfounderallele)
            # exchange: one all-gather per maximal run of consecutive indices in the widest block
            # (nothing at all when no rank has anything to share: every rank knows the whole plan)
            runs = [contiguous_runs(b) for b in blocks]
            for k in range(max(len(r) for r in runs)):
                self._exchange(runs, k)

    def gather(self, individuals):
        """All-gather the haplostructs of the listed individuals from their owners to every rank
        (lineage partition: e.g. the last generation, for a check or a joint emission)."""
        individuals = np.asarray(individuals, np.int64)
        if self.owner is None:
            return
        lists = [individuals[self.owner[individuals] == r] for r in range(self.world)]
        runs = [contiguous_runs(np.sort(b)) for b in lists]
        for k in range(max(len(r) for r in runs)):
            self._exchange(runs, k)

    def _exchange(self, runs, k):
        if k < len(runs[self.rank]):
            first, count = runs[self.rank][k]
            off, fo, po = self.shard.export_range(first, count)
        else:   # nothing to contribute in this round: an empty block
            dev = "cuda" if torch.cuda.is_available() and dist.get_backend(self.group) == "nccl" else "cpu"
            off = torch.zeros(1, dtype=torch.int64, device=dev)
            fo = torch.zeros(0, dtype=torch.int32, device=dev)
            po = torch.zeros(0, dtype=torch.float64, device=dev)
        got = allgather_csr(off, fo, po, self.group)
        self.collectives += 1
        for r in range(self.world):
            if r == self.rank or k >= len(runs[r]):
                continue
            first, count = runs[r][k]
            self.shard.import_range(first, count, *got[r])
            self.exchanged_segments += int(got[r][1].numel())
