"""Host-target emission timing (bench e2e leg in isolation)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
h_founder = torch.empty((L, N * P), dtype=torch.uint8, pin_memory=True)
h_dose = torch.empty((L, N), dtype=torch.uint8, pin_memory=True)
ht = ps.EmitTargets()
ht.founder, ht.founder_pitch, ht.founder_bytes = h_founder.data_ptr(), N * P, 1
ht.dose, ht.dose_pitch, ht.dose_which = h_dose.data_ptr(), N, ps.DOSE_MAX_ALLELE
ht.memory = ps.MEM_HOST
full = "--full" in sys.argv
if "--stream" in sys.argv:
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
for k in range(4):
    ts = [time.perf_counter()]
    if full:
        ctx.set_pedigree(w["p0"], w["p1"])
    ts.append(time.perf_counter())
    ctx.reset(bench.SEED, k + 1)
    ts.append(time.perf_counter())
    if full:
        ctx.set_map(w["off"], w["pos"])
    ts.append(time.perf_counter())
    if full:
        ctx.set_allele_codes(w["code"])
    ts.append(time.perf_counter())
    ctx.simulate()
    ts.append(time.perf_counter())
    print("set_pedigree %.2f reset %.2f set_map %.2f set_codes %.2f simulate %.2f ms" % tuple((b - a) * 1e3 for a, b in zip(ts, ts[1:])))
    t0 = time.perf_counter()
    ctx.emit_into(ht, 0, N, 0, L)
    t1 = time.perf_counter()
    st = ctx.stats()
    print("emit host %.2f ms, device total %.2f ms, kernel %.3f ms, launches %d" % ((t1 - t0) * 1e3, st["ms_emit_total"], st["ms_emit_kernel"], st["emit_kernel_launches"]))
