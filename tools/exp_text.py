"""Throughput of psim_emit_text (device-formatted table lines into page-locked host buffers) on bench configs[1]."""
import ctypes as C
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import pedigreesim_b200 as ps

w = bench.workload(10000)
P, L, N = w["P"], w["L"], w["N"]
ctx = ps.PsimContext(P, seed=bench.SEED, **w["model"])
ctx.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
ctx.set_pedigree(w["p0"], w["p1"])
ctx.set_map(w["off"], w["pos"])
ctx.set_allele_codes(w["code"])
ctx.set_locus_names(["m%02d_%05d" % (l // 2000, l % 2000) for l in range(L)])
ctx.set_allele_names([["0", "1"]] * L)
ctx.simulate()
blk = 2000
cap_f = blk * (16 + N * P * 2 + 1)
cap_d = blk * (16 + N * 2 + 1)
bufs = {k: torch.empty(c, dtype=torch.uint8, pin_memory=True) for k, c in (("founder", cap_f), ("allele", cap_f), ("dose", cap_d))}
for rep in range(3):
    t0 = time.perf_counter()
    total = 0
    for l0 in range(0, L, blk):
        t = ps.TextTargets()
        for k, b in bufs.items():
            setattr(t, k, b.data_ptr())
            setattr(t, k + "_capacity", b.numel())
        t.dose_which = ps.DOSE_MAX_ALLELE
        ctx._chk(ctx.L.psim_emit_text(ctx.h, 0, N, l0, blk, C.byref(t)))
        total += t.founder_bytes + t.allele_bytes + t.dose_bytes
    dt = time.perf_counter() - t0
    print("emit_text: %.2f GB of table text (3 tables, %d loci x %d individuals) in %.3f s = %.1f GB/s, %.1f G allele-loci/s"
          % (total / 1e9, L, N, dt, total / dt / 1e9, N * L * P / dt / 1e9))
