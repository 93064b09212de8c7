"""One large pedigree over several GPUs.

One process per GPU. Every rank holds the whole pedigree in its own libpsim context but simulates
only its block of each generation; after a generation the new HaploStructs are exchanged with one
variable-length all-gather (sizes, then the packed CSR) so that any rank can read any parent in the
next generation. Replicate pedigrees need none of this (no data-path collective: each rank just
uses its own replicate ids).

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

    def __init__(self, shard, parent0, parent1, group=None):
        self.shard = shard
        self.p0 = np.asarray(parent0, np.int32)
        self.p1 = np.asarray(parent1, np.int32)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.exchanged_segments = 0
        self._levels = None   # generations of the whole pedigree (has_hs None): computed once, the pedigree is fixed

    def simulate(self, maxfounderallele=-1, has_hs=None):
        if has_hs is None:
            if self._levels is None:
                self._levels = levelize(self.p0, self.p1)
            levels = self._levels
        else:
            levels = levelize(self.p0, self.p1, has_hs)
        # founders are initialised by every rank (deterministic, no exchange needed)
        self.shard.simulate_subset(np.zeros(0, np.int32), maxfounderallele)
        for inds in levels:
            blocks = block_partition(inds, self.world)
            mine = blocks[self.rank]
            if len(mine):
                self.shard.simulate_subset(mine, maxfounderallele)
            # exchange: one all-gather per maximal run of consecutive indices in the widest block
            runs = [contiguous_runs(b) for b in blocks]
            for k in range(max(len(r) for r in runs)):
                if k < len(runs[self.rank]):
                    first, count = runs[self.rank][k]
                    off, fo, po = self.shard.export_range(first, count)
                else:   # nothing to contribute in this round: an empty block
                    dev = "cuda" if torch.cuda.is_available() and dist.get_backend(self.group) == "nccl" else "cpu"
                    off = torch.zeros(1, dtype=torch.int64, device=dev)
                    fo = torch.zeros(0, dtype=torch.int32, device=dev)
                    po = torch.zeros(0, dtype=torch.float64, device=dev)
                got = allgather_csr(off, fo, po, self.group)
                for r in range(self.world):
                    if r == self.rank or k >= len(runs[r]):
                        continue
                    first, count = runs[r][k]
                    self.shard.import_range(first, count, *got[r])
                    self.exchanged_segments += int(got[r][1].numel())
