#!/usr/bin/env python
"""Short profiling driver: the bench workload (configs[1]) for a few device-resident steps, no
host-side legs. Used under ncu (launch list / --set full capture of k_emit_tiles)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import pedigreesim_b200 as ps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--popsize", type=int, default=10000)
    ap.add_argument("--alleles", action="store_true", help="also emit the allele-code table")
    ap.add_argument("--no-dose", action="store_true")
    ap.add_argument("--no-founder", action="store_true")
    ap.add_argument("--no-codes", action="store_true", help="no founder genotypes: dose of founder allele 0")
    ap.add_argument("--packed", action="store_true", help="PSIM_EMIT_PACKED4 tables")
    ap.add_argument("--replicates", type=int, default=1, help="replicate populations side by side in the context and its tables")
    ap.add_argument("--scale", type=float, default=1.0, help="scale chromosome lengths (1e-4: no crossovers at all)")
    args = ap.parse_args()
    w = bench.workload(args.popsize)
    for k in ("lengths", "centro", "pos"):
        w[k] = w[k] * args.scale
    P, L, N = w["P"], w["L"], w["N"]
    R = args.replicates
    ctx = ps.PsimContext(P, seed=bench.SEED, replicate_count=R, **w["model"])
    ctx.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
    ctx.set_pedigree(w["p0"], w["p1"])
    ctx.set_map(w["off"], w["pos"])
    if not args.no_codes:
        ctx.set_allele_codes(w["code"])
    fp = (R * N * P // (2 if args.packed else 1) + 127) // 128 * 128
    dp = ((R * N + 1) // (2 if args.packed else 1) + 127) // 128 * 128
    d_founder = torch.empty((L, fp), dtype=torch.uint8, device="cuda")
    d_dose = torch.empty((L, dp), dtype=torch.uint8, device="cuda")
    d_allele = torch.empty((L, fp), dtype=torch.uint8, device="cuda") if args.alleles else None
    tgt = ps.EmitTargets()
    if not args.no_founder:
        tgt.founder, tgt.founder_pitch, tgt.founder_bytes = d_founder.data_ptr(), fp, 1
    if not args.no_dose:
        tgt.dose, tgt.dose_pitch, tgt.dose_which = d_dose.data_ptr(), dp, (0 if args.no_codes else ps.DOSE_MAX_ALLELE)
    if args.alleles:
        tgt.allele, tgt.allele_pitch = d_allele.data_ptr(), fp
    tgt.memory = ps.MEM_DEVICE
    tgt.flags = ps.EMIT_PACKED4 if args.packed else 0
    div = 2 if args.packed else 1
    for k in range(args.steps):
        t0 = time.perf_counter()
        ctx.reset(bench.SEED, k + 1)
        t1 = time.perf_counter()
        ctx.simulate()
        t2 = time.perf_counter()
        ctx.emit_into(tgt, 0, N, 0, L)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        st = ctx.stats()
        print("step %d: reset %.2f ms, simulate %.2f ms (device %.2f), emit %.2f ms (kernel %.3f ms = %.0f GB/s)" % (
            k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["ms_simulate"], (t3 - t2) * 1e3, st["ms_emit_kernel"],
            R * N * L * ((0 if args.no_founder else P) + (0 if args.no_dose else 1) + (P if args.alleles else 0)) / div / st["ms_emit_kernel"] / 1e6))
    ctx.close()


if __name__ == "__main__":
    main()
