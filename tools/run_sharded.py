#!/usr/bin/env python
"""One large pedigree sharded over the GPUs of a box (BASELINE.json configs[3] at reduced size):
diploid RIL chain F2 -> F<g> by selfing, individuals of each generation split over the ranks, the new
HaploStructs that another rank needs exchanged after every generation. --impl library
(default) is psim_simulate_sharded: the generation loop and the NCCL exchange inside libpsim; --impl
python is the same partition driven from pedigreesim_b200/sharded.py over torch.distributed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_sharded.py --lines 200000 --generations 8

Every rank ends with the whole pedigree; rank 0 also simulates it unsharded and checks equality
(Philox streams are keyed by individual, so the sharding must not change a bit).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pedigreesim_b200 as ps  # noqa: E402
from pedigreesim_b200 import sharded  # noqa: E402


def ril_pedigree(lines, generations):
    p0, p1 = [-1, -1, 0], [-1, -1, 1]
    prev = np.full(lines, 2, np.int32)
    for _ in range(generations - 1):
        first = len(p0)
        p0 += prev.tolist()
        p1 += prev.tolist()
        prev = np.arange(first, first + lines, dtype=np.int32)
    return np.array(p0, np.int32), np.array(p1, np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=100000)
    ap.add_argument("--generations", type=int, default=8)
    ap.add_argument("--chromosomes", type=int, default=10)
    ap.add_argument("--check", type=int, nargs="?", const=1, default=1)
    ap.add_argument("--impl", default="library", choices=["library", "python"])
    ap.add_argument("--out", default=None, help="append the JSON line to this file as well")
    ap.add_argument("--partition", default="block", choices=["block", "lineage"],
                    help="lineage: a rank keeps its lines, only individuals with a child on another rank are shared")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p0, p1 = ril_pedigree(args.lines, args.generations)
    lengths = [1.5] * args.chromosomes
    centro = [0.75] * args.chromosomes
    ctx = ps.PsimContext(2, seed=12345, device=local)
    ctx.set_chromosomes(lengths, centro)
    ctx.set_pedigree(p0, p1)
    part = ps.PARTITION_BLOCK if args.partition == "block" else ps.PARTITION_LINEAGE
    if args.impl == "library":
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(ps.shard_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.shard_init(rank, world, uid.cpu().numpy().tobytes())
        run = lambda: ctx.simulate_sharded(partition=part)   # noqa: E731
    else:
        drv = sharded.ShardedPedigree(sharded.PsimShard(ctx), p0, p1, partition=args.partition)
        run = drv.simulate
    # first pass: the slab and the exchange buffers grow to their final size (cudaMalloc / cudaFree of
    # gigabytes); the timed pass re-simulates the same pedigree after psim_reset into that memory
    run()
    torch.cuda.synchronize()
    ctx.reset(12345, 0)
    if args.impl == "python":
        drv.exchanged_segments = 0
        drv.collectives = 0
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    n = len(p0)
    ok = True
    unsharded_dt = None
    if args.impl == "python":
        exchanged, collectives = drv.exchanged_segments, drv.collectives
    else:
        st = ctx.shard_stats()
        exchanged, collectives = st["segments_received"], st["collectives"]
    if args.check:
        last = n - args.lines
        if args.partition == "lineage":   # outside the timed region: bring the last generation together for the check
            if args.impl == "python":
                drv.gather(np.arange(last, n))
            else:
                ctx.shard_gather(np.arange(last, n, dtype=np.int32))
        so, fo, po = ctx.get_haplostructs(first=last, count=args.lines)
        if rank == 0:
            ref = ps.PsimContext(2, seed=12345, device=local)
            ref.set_chromosomes(lengths, centro)
            ref.set_pedigree(p0, p1)
            ref.simulate()
            ref.reset(12345, 0)          # warm pass like the sharded run, then the timed one
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ref.simulate()
            torch.cuda.synchronize()
            unsharded_dt = time.perf_counter() - t1
            s2, f2, x2 = ref.get_haplostructs(first=last, count=args.lines)
            ok = bool(np.array_equal(so, s2) and np.array_equal(fo, f2) and np.array_equal(po, x2))
            if not ok:
                print("mismatch vs unsharded: seg_off %s founder %s pos %s (sizes %d / %d)" % (
                    np.array_equal(so, s2), len(fo) == len(f2) and np.array_equal(fo, f2),
                    len(po) == len(x2) and np.array_equal(po, x2), len(fo), len(f2)), flush=True)
            ref.close()
        digest = torch.tensor([float(fo.sum() % 1000003), float(len(fo))], device="cuda", dtype=torch.float64)
        all_d = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(all_d, digest)
        same = all(bool(torch.equal(all_d[0], d)) for d in all_d)
        if not same and rank == 0:
            print("rank digests differ:", [d.tolist() for d in all_d], flush=True)
        ok = ok and same
    if rank == 0:
        meioses = 2 * (n - 2)
        line = json.dumps({"n_gpus": world, "world": world, "impl": args.impl, "transport": "nccl", "individuals": n, "meioses": meioses,
                           "seconds": dt, "meioses_per_s": meioses / dt,
                           "partition": args.partition, "exchanged_segments_per_rank": exchanged, "collectives": collectives,
                           "bit_identical_to_unsharded": ok, "identical": ok,
                           "unsharded_seconds_one_gpu": unsharded_dt,
                           "speedup_vs_one_gpu": None if not unsharded_dt else unsharded_dt / dt})
        print(line)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")
    ctx.close()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
