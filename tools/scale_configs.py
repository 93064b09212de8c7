#!/usr/bin/env python
"""BASELINE.json configs[2..4] at their full individual counts on one B200 (synthetic inputs of
SURVEY.md §8(d)): simulate the whole pedigree, emit blocks of loci into a reused device buffer, time
both, and cross-check the row-image emission against the tiled kernel on the same block
(two independent kernels must agree byte for byte) plus size-independent properties.
One JSON line per config (kept in profiles/scale_r1.jsonl).

    python tools/scale_configs.py [3] [4] [5]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import _cases  # noqa: E402
import pedigreesim_b200 as ps  # noqa: E402


def emit_block(ctx, P, n_ind_first, n_ind, l0, nl, founder, dose, which, path=0):
    fp, dp = founder.shape[1], dose.shape[1]
    t = ps.EmitTargets()
    t.flags = ps.emit_path(path)
    t.founder, t.founder_pitch, t.founder_bytes = founder.data_ptr(), fp, 1
    t.dose, t.dose_pitch, t.dose_which = dose.data_ptr(), dp, which
    t.memory = ps.MEM_DEVICE
    ctx.emit_into(t, n_ind_first, n_ind, l0, nl)
    return ctx.stats()


def run(name, P, p0, p1, lengths, centro, pref, fq, loci_per_chrom, emit_first, emit_count, block_loci, n_blocks, model, replicate_count=1):
    C = len(lengths)
    out = {"config": name, "ploidy": P, "individuals": int(len(p0)), "chromosomes": C, "loci": int(C * loci_per_chrom),
           "replicates": replicate_count}
    t0 = time.perf_counter()
    ctx = ps.PsimContext(P, seed=12345, replicate_count=replicate_count, **model)
    ctx.set_chromosomes(lengths, centro, pref, fq)
    ctx.set_pedigree(p0, p1)
    out["setup_s"] = round(time.perf_counter() - t0, 3)
    ctx.simulate()          # cold: kernels are loaded on first use, scratch and slab are allocated
    ctx.reset(12345, 0)
    before = ctx.stats()
    t0 = time.perf_counter()
    ctx.simulate()
    torch.cuda.synchronize()
    st = ctx.stats()
    n_mei, n_units = st["n_meioses"] - before["n_meioses"], st["n_units"] - before["n_units"]   # (the counters run on across calls)
    out.update(simulate_s=round(time.perf_counter() - t0, 3), simulate_device_ms=round(st["ms_simulate"], 2), meioses=int(n_mei),
               units=int(n_units), segments=int(st["n_segments"]),
               meioses_per_s=round(n_mei / (st["ms_simulate"] * 1e-3)))
    off, pos = _cases.uniform_map(lengths, [loci_per_chrom] * C)
    ctx.set_map(off, pos)
    R = replicate_count
    cols = R * emit_count * P
    fp = (cols + 127) // 128 * 128
    dp = (R * emit_count + 127) // 128 * 128
    founder = torch.empty((block_loci, fp), dtype=torch.uint8, device="cuda")
    dose = torch.empty((block_loci, dp), dtype=torch.uint8, device="cuda")
    L = int(off[-1])
    starts = np.linspace(0, L - block_loci, n_blocks).astype(np.int64) if L > block_loci else np.array([0])
    block_loci = min(block_loci, L)
    ms_kernel, ms_total = [], []
    for k, l0 in enumerate(starts):
        for rep in range(2 if k == 0 else 1):   # first block twice: warm-up (index runs, allocations)
            st = emit_block(ctx, P, emit_first, emit_count, int(l0), block_loci, founder, dose, 0)
        ms_kernel.append(st["ms_emit_kernel"])
        ms_total.append(st["ms_emit_total"])
    algo = block_loci * R * emit_count * (P + 1)
    out.update(emit_block_loci=int(block_loci), emit_blocks=len(starts), emit_bytes_per_block=int(algo),
               emit_kernel_ms=round(float(np.mean(ms_kernel)), 3), emit_all_kernels_ms=round(float(np.mean(ms_total)), 3),
               emit_kernel_GBps=round(algo / np.mean(ms_kernel) / 1e6, 1),
               emit_G_allele_loci_per_s=round(block_loci * cols / np.mean(ms_total) / 1e6, 1))
    # cross-check the last block against the tiled kernel and against invariants
    f_rows = founder[:, :cols].clone()
    d_rows = dose[:, :R * emit_count].clone()
    founder.zero_()
    dose.zero_()
    torch.cuda.synchronize()   # torch's stream is not the library's
    emit_block(ctx, P, emit_first, emit_count, int(starts[-1]), block_loci, founder, dose, 0, path=ps.PATH_TILES)
    same = bool(torch.equal(f_rows, founder[:, :cols]) and torch.equal(d_rows, dose[:, :R * emit_count]))
    n_found = int((p0 < 0).sum())
    in_range = bool(int(f_rows.max()) < n_found * P)
    dose_ok = bool(torch.equal((f_rows.view(block_loci, -1, P) == 0).sum(dim=2).to(torch.uint8), d_rows))
    out.update(rows_equal_tiled_kernel=same, founder_alleles_in_range=in_range, dose_equals_count_of_allele_0=dose_ok,
               gpu_mem_used_GB=round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9, 1))
    ctx.close()
    print(json.dumps(out), flush=True)
    assert same and in_range and dose_ok


def main():
    which = [int(a) for a in sys.argv[1:]] or [3, 4, 5]
    if 3 in which:   # hexaploid F1 of 100 000, 21 chromosomes x 10 000 loci, natural pairing
        p0, p1 = _cases.std_pedigree("F1", 100000)
        run("configs[2] hexaploid F1 100000, 21 x 10000 loci", 6, p0, p1, [1.0] * 21, [0.5] * 21, [0.0] * 21, [0.0] * 21, 10000,
            0, len(p0), 4000, 4, dict(natural_pairing=True))
    if 4 in which:   # diploid RIL chain F2 -> F8, 1 M lines, 10 chromosomes x 50 000 loci; emit F8 only
        lines, gens = 1000000, 8
        p0, p1 = _cases.ril_pedigree(lines, gens)
        n = len(p0)
        run("configs[3] diploid RIL F2->F8, 1M lines, 10 x 50000 loci (F8 emitted)", 2, p0, p1, [1.5] * 10, [0.75] * 10, None, None, 50000,
            n - lines, lines, 2000, 4, dict())
    if 5 in which:   # replicate tetraploid pedigrees: F1 of 200, 5 chromosomes x 100 loci, natural pairing; 2000 replicates per context
        p0, p1 = _cases.std_pedigree("F1", 200)
        run("configs[4] 2000 of the 100000 replicate tetraploid F1(200) pedigrees, 5 x 100 loci", 4, p0, p1, [1.0] * 5, [0.5] * 5,
            [0.25] * 5, [0.0] * 5, 100, 0, len(p0), 500, 1, dict(natural_pairing=True), replicate_count=2000)


if __name__ == "__main__":
    main()
