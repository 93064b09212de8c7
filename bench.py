#!/usr/bin/env python
"""Benchmark of the PedigreeSim hot path on B200: meioses/s and genotype calls/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic input: R (--replicates, default 8) replicate
populations of BASELINE.json configs[1] (autotetraploid F1 of 10 000, quadrivalent fraction 0.5, preferential pairing
0.3, 12 chromosomes x 2 000 loci, biallelic founder genotypes) side by side in one context - the batch of a Monte Carlo
sweep, the job replicate sharding exists for: all R x 20 000 meioses (K1), the breakpoint -> locus-run index (K0) and
the dense emission of the founder-allele matrix plus the MAX-allele dose table of all R populations in ONE launch (K2,
the HBM-write-bound kernel; R*N*L*(P+1) algorithmic bytes). With N > 1 GPUs every rank simulates and emits its own
replicate populations (weak scaling, no data-path collective). `single_population` reports R = 1 (one population on
its own, not pipelined); --replicates 1 runs the whole line that way (round 1's configuration).

Prints ONE JSON line (rank 0). `value` is measured with inputs resident in HBM and outputs written to HBM (u8 tables):
three contexts in rotation, the next batch's meioses (psim_simulate_async) run while the current one is emitted
(PSIM_EMIT_ASYNC), one host wait per step. `e2e` is the same metric through the C ABI with host buffers, one population
per step (pedigree / map / allele-code uploads and the device->host copy of both tables inside the timed region), with
PSIM_EMIT_PACKED4 tables (two 4-bit cells per byte: the 8 founder alleles and the doses 0..4 of this workload fit;
SURVEY.md 8(d) counts a packed format at its packed size); `e2e.u8` is the same leg with one byte per cell.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 12345
CPU_SAMPLE_POPSIZE = 1000   # offspring per CPU population: the one sample size of cpu_baseline, --impl reference and DESIGN.md


# ------------------------------------------------------------------------------------ workload
def workload(popsize=10000):
    """BASELINE.json configs[1] as SURVEY.md §8(d) item 2 spells it out."""
    P, C, loci = 4, 12, 2000
    lengths = np.array([(80 + 10 * (c % 5)) / 100.0 for c in range(C)])
    centro = 0.4 * lengths
    pref = np.full(C, 0.3)
    fq = np.full(C, 0.5)
    p0 = np.array([-1, -1] + [0] * popsize, np.int32)  # POPTYPE=F1 (Main.java:1471-1476)
    p1 = np.array([-1, -1] + [1] * popsize, np.int32)
    off = np.arange(C + 1, dtype=np.int64) * loci
    pos = np.concatenate([ln * (np.arange(loci) + 0.5) / loci for ln in lengths])
    A = 2 * P
    l = np.arange(C * loci, dtype=np.uint64)[:, None]
    a = np.arange(A, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = (l * np.uint64(0x9E3779B97F4A7C15) + a * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(29)
    code = (h & np.uint64(1)).astype(np.uint8)
    code[code.max(axis=1) == 0, 0] = 1  # keep every locus biallelic so the codes are ranks 0/1
    code[code.min(axis=1) == 1, 0] = 0
    return dict(P=P, C=C, L=C * loci, N=len(p0), M=2 * popsize, lengths=lengths, centro=centro, pref=pref, fq=fq,
                p0=p0, p1=p1, off=off, pos=pos, code=code, A=A,
                model=dict(kosambi=False, natural_pairing=False, allow_no_recomb=True))


WORKLOAD_NAME = ("configs[1]: autotetraploid F1 of 10000 (POPTYPE=F1), NATURALPAIRING=0, quadrivalents 0.5, "
                 "prefPairing 0.3, 12 chromosomes (80-120 cM) x 2000 loci, biallelic .gen")


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks and throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Number of samples so far (start of a window)."""
        return len(self.lines)

    def wait_for(self, n, timeout):
        t0 = time.time()
        while self.proc and len(self.lines) < n and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(w, offspring, steps, warmup):
    """The oracle (CPU restatement of the reference, java.util.Random draw order) on `offspring`
    F1 individuals of the same workload: simulate + founder table + MAX-allele dose table."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    ws = workload(offspring)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        o = _oracle.Oracle(ws["P"], rng=_oracle.RNG_JAVA, seed=SEED + s, **ws["model"])
        o.set_chromosomes(ws["lengths"], ws["centro"], ws["pref"], ws["fq"])
        o.set_pedigree(ws["p0"], ws["p1"])
        o.simulate()
        o.set_map(ws["off"], ws["pos"])
        o.set_allele_codes(ws["code"])
        o.emit(_oracle.EMIT_FOUNDER)
        o.emit(_oracle.EMIT_DOSE, which=_oracle.MAX_ALLELE)
        o.close()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    cells = ws["N"] * ws["L"] * ws["P"]
    return cells / t / 1e9, t, ws


def _reference_worker(job):
    offspring, steps, warmup, seed_shift = job
    global SEED
    SEED = SEED + 1000 * seed_shift   # every replicate population has its own seed, like jar SEED=seed+r
    v, t, ws = cpu_port_run(workload(), offspring, steps, warmup)
    return v, t, ws["M"]


def find_jar():
    """The reference's own jar and a JVM to run it, if this box has both. The jar is copied (unmodified) to
    baseline/_ref/ by __graft_entry__.build() where /root/reference exists; no JVM ships in this image."""
    import shutil
    java = shutil.which("java")
    jar = os.path.join(ROOT, "baseline", "_ref", "PedigreeSim.jar")
    return (java, jar) if java and os.path.exists(jar) else (None, None)


def _jar_worker(job):
    """One population = one JVM run of PedigreeSim.jar on its own .par (SEED = 12345 + replicate)."""
    java, jar, offspring, steps, warmup, k = job
    import tempfile
    times = []
    for s in range(warmup + steps):
        with tempfile.TemporaryDirectory() as d:
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_config.py"), "2", d, "--popsize", str(offspring)],
                                  stdout=subprocess.DEVNULL)
            par = os.path.join(d, "x.par")
            txt = open(par).read().replace("SEED = 12345", "SEED = %d" % (SEED + 1000 * k + s))
            open(par, "w").write(txt)
            t0 = time.perf_counter()
            subprocess.check_call([java, "-jar", jar, "x.par"], cwd=d, stdout=subprocess.DEVNULL)
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    ws = workload(offspring)
    t = float(np.mean(times))
    return ws["N"] * ws["L"] * ws["P"] / t / 1e9, t, ws["M"]


def run_reference(args):
    """The reference's CPU path on this job. One population is one single-threaded run (the jar shares
    one java.util.Random and cannot be threaded), but the job is a sequence of independent replicate
    populations (one per step and rank), so the host runs one population per core side by side on
    every core it has - the way PedigreeSim is run for a Monte Carlo sweep (one JVM per core).
    With a JVM on the box this is PedigreeSim.jar itself (kind "jar", file output included, as the
    reference has no other sink); otherwise the oracle port of the same path (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload()
    try:
        n_cpu = len(os.sched_getaffinity(0))
    except AttributeError:
        n_cpu = os.cpu_count() or 1
    procs = max(1, n_cpu)
    offspring = CPU_SAMPLE_POPSIZE
    java, jar = find_jar()
    import multiprocessing as mp
    if java:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_jar_worker, [(java, jar, offspring, args.steps, args.warmup, k) for k in range(procs)])
        kind, note = "jar", "PedigreeSim.jar, one JVM per core, its own output files written"
    elif procs == 1:
        v1, t1, ws = cpu_port_run(w, offspring, args.steps, args.warmup)
        res = [(v1, t1, ws["M"])]
        kind, note = "port", None
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_reference_worker, [(offspring, args.steps, args.warmup, k) for k in range(procs)])
        kind, note = "port", None
    if kind == "port":
        note = ("no JVM on this box: the reference jar cannot run; this is the oracle restatement (one thread per "
                "population like the jar), without the jar's text formatting")
    t = max(r[1] for r in res)                       # the slowest population closes the step
    v = sum(r[0] * r[1] for r in res) / t            # cells of all populations / that time
    meioses_per_s = sum(r[2] for r in res) / t
    sample = "%d replicate population(s) side by side, each the same workload with POPSIZE=%d of 10000 offspring per step (all 24000 loci)" % (procs, offspring)
    out = {
        "impl": "reference", "metric": "genotype calls/sec (G allele-loci/s)", "value": v, "unit": "G allele-loci/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "G allele-loci/s", "cores": procs, "kind": kind, "sample": sample, "note": note},
        "e2e": {"value": v, "unit": "G allele-loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "meioses_per_s": meioses_per_s,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(index):
    """Run this rank on the CPU cores next to its GPU (NVML's affinity mask), so that the page-locked
    host tables of the end-to-end leg are first touched on the NUMA node the GPU's PCIe link hangs
    off. Best effort: silently skipped when NVML or sched_setaffinity is not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def replicate_sweep(ps, torch, local, rank, R=2000):
    """BASELINE.json configs[4] on this GPU: R independent replicate tetraploid pedigrees (F1 of 200, natural pairing,
    prefPairing 0.25, 5 chromosomes x 100 loci) side by side in one context; simulate, then emit the founder-allele and
    founder-allele-0 dose tables of all replicates (u8, device). Reports the emission against the HBM roofline end to
    end (every emission kernel, not only the streaming one)."""
    P, C, n, loci = 4, 5, 200, 100
    p0 = np.array([-1, -1] + [0] * n, np.int32)
    p1 = np.array([-1, -1] + [1] * n, np.int32)
    N = n + 2
    off = np.arange(C + 1, dtype=np.int64) * loci
    pos = np.concatenate([(np.arange(loci) + 0.5) / loci for _ in range(C)])
    ctx = ps.PsimContext(P, seed=SEED, replicate_first=1000000 + rank * R, replicate_count=R, device=local, natural_pairing=True)
    ctx.set_chromosomes([1.0] * C, [0.5] * C, [0.25] * C, [0.0] * C)
    ctx.set_pedigree(p0, p1)
    ctx.set_map(off, pos)
    L = C * loci
    cols = R * N * P
    fp, dp = (cols + 127) // 128 * 128, (R * N + 127) // 128 * 128
    f = torch.empty((L, fp), dtype=torch.uint8, device="cuda")
    d = torch.empty((L, dp), dtype=torch.uint8, device="cuda")
    t = ps.EmitTargets()
    t.founder, t.founder_pitch, t.founder_bytes = f.data_ptr(), fp, 1
    t.dose, t.dose_pitch, t.dose_which = d.data_ptr(), dp, 0
    t.memory = ps.MEM_DEVICE
    sim, ek, et = [], [], []
    for k in range(4):
        ctx.reset(SEED + k, 1000000 + rank * R)
        ctx.simulate()
        ctx.emit_into(t, 0, N, 0, L)
        st = ctx.stats()
        if k:
            sim.append(st["ms_simulate"]); ek.append(st["ms_emit_kernel"]); et.append(st["ms_emit_total"])
    ctx.close()
    bytes_ = L * R * N * (P + 1)
    peak, _ = measured_peak()
    ms_sim, ms_k, ms_all = float(np.mean(sim)), float(np.mean(ek)), float(np.mean(et))
    return {"workload": "configs[4]: %d replicate tetraploid F1(200) pedigrees in one context, 5 chromosomes x 100 loci" % R,
            "meioses_per_s": R * 2 * n / (ms_sim * 1e-3), "simulate_ms": ms_sim,
            "emit_kernel_ms": ms_k, "emit_all_kernels_ms": ms_all, "bytes": bytes_,
            "G_allele_loci_per_s": L * cols / (ms_all * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "scope": "all emission kernels of the call (index, band plan, rows)", "achieved": bytes_ / (ms_all * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": bytes_ / (ms_all * 1e-3) / 1e9 / peak}}


def sharded_block(ps, torch, dist, rank, world, local, lines, generations=8, chromosomes=10):
    """BASELINE.json configs[3]: ONE pedigree over the GPUs of the box. Diploid RIL chain P1 x P2 -> F1 -> `lines` F2 plants
    selfed to F<generations>, 10 chromosomes of 150 cM: psim_simulate_sharded (generation loop and NCCL exchange inside
    libpsim) with both partitions, against psim_simulate of the whole pedigree on one GPU (rank 0's). Every timing is a
    second pass after psim_reset (slab and exchange buffers at their final size), max over ranks; the last generation is
    compared bit for bit with the one-GPU run on rank 0."""
    p0 = np.empty(3 + lines * (generations - 1), np.int32)
    p0[:3] = (-1, -1, 0)
    p1 = p0.copy()
    p1[2] = 1
    p0[3:3 + lines] = 2
    p0[3 + lines:] = np.arange(3, 3 + lines * (generations - 2), dtype=np.int32)
    p1[3:] = p0[3:]
    n = len(p0)
    lengths, centro = [1.5] * chromosomes, [0.75] * chromosomes
    ctx = ps.PsimContext(2, seed=SEED, device=local)
    ctx.set_chromosomes(lengths, centro)
    ctx.set_pedigree(p0, p1)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(ps.shard_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.shard_init(rank, world, uid.cpu().numpy().tobytes())
    sample_first, sample_count = n - 4096, 4096
    ref = None
    out = {"workload": "configs[3]: diploid RIL chain, %d F2 lines selfed to F%d = %d individuals, %d chromosomes x 150 cM, one pedigree over %d GPUs"
                       % (lines, generations, n, chromosomes, world), "meioses": 2 * (n - 2)}
    if rank == 0:
        ctx.simulate()
        ctx.reset(SEED, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.simulate()
        torch.cuda.synchronize()
        out["one_gpu_seconds"] = time.perf_counter() - t0
        ref = ctx.get_haplostructs(sample_first, sample_count)
        out["segments"] = int(ctx.stats()["n_segments"])
    for name, part in (("block", ps.PARTITION_BLOCK), ("lineage", ps.PARTITION_LINEAGE)):
        ctx.reset(SEED, 0)
        ctx.simulate_sharded(partition=part)
        ctx.reset(SEED, 0)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.simulate_sharded(partition=part)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = ctx.shard_stats()
        if part == ps.PARTITION_LINEAGE:   # outside the timed region: bring the sample together for the check
            ctx.shard_gather(np.arange(sample_first, sample_first + sample_count, dtype=np.int32))
        same = None
        if rank == 0:
            got = ctx.get_haplostructs(sample_first, sample_count)
            same = all(np.array_equal(a, b) for a, b in zip(ref, got))
        out[name] = {"seconds": float(t[0]), "meioses_per_s": 2 * (n - 2) / float(t[0]),
                     "speedup_vs_one_gpu": None if rank else out["one_gpu_seconds"] / float(t[0]),
                     "segments_received_rank0": int(st["segments_received"]), "bytes_sent_rank0": int(st["bytes_sent"]),
                     "exchange_steps": int(st["collectives"]), "sample_bit_identical_to_one_gpu": same}
    ctx.close()
    return out


def replicate_batch(ps, torch, local, rank, w, R=8):
    """The bench workload with R replicate populations side by side in one context (replicate_count = R): R = 8 is what a
    Monte Carlo sweep over configs[1] looks like to the library, R = 1 one population on its own. One population's table
    row is 40 KB; R = 8 makes it 320 KB, and the row-image kernel then runs at the pace of HBM instead of at the pace a
    launch of two units per SM allows. Not pipelined: simulate and emit alone, three passes."""
    P, L, N = w["P"], w["L"], w["N"]
    ctx = ps.PsimContext(P, seed=SEED, replicate_first=500000 + rank * R, replicate_count=R, device=local, **w["model"])
    ctx.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
    ctx.set_pedigree(w["p0"], w["p1"])
    ctx.set_map(w["off"], w["pos"])
    ctx.set_allele_codes(w["code"])
    fp, dp = (R * N * P + 127) // 128 * 128, (R * N + 127) // 128 * 128
    f = torch.empty((L, fp), dtype=torch.uint8, device="cuda")
    d = torch.empty((L, dp), dtype=torch.uint8, device="cuda")
    t = ps.EmitTargets()
    t.founder, t.founder_pitch, t.founder_bytes = f.data_ptr(), fp, 1
    t.dose, t.dose_pitch, t.dose_which = d.data_ptr(), dp, ps.DOSE_MAX_ALLELE
    t.memory = ps.MEM_DEVICE
    sim, ek, et, launches = [], [], [], 0
    for k in range(4):
        ctx.reset(SEED + k, 500000 + rank * R)
        ctx.simulate()
        ctx.emit_into(t, 0, N, 0, L)
        st = ctx.stats()
        if k:
            sim.append(st["ms_simulate"]); ek.append(st["ms_emit_kernel"]); et.append(st["ms_emit_total"]); launches += st["emit_kernel_launches"]
    ctx.close()
    peak, _ = measured_peak()
    bytes_ = R * N * L * (P + 1)
    k_ms = float(np.sum(ek)) / max(1, launches)
    ms_sim, ms_all = float(np.mean(sim)), float(np.mean(et))
    return {"workload": "%d replicate populations of the bench workload side by side in one context (rows of %d KB)" % (R, R * N * P // 1000),
            "simulate_ms": ms_sim, "emit_kernel_ms": k_ms, "emit_all_kernels_ms": ms_all, "bytes": bytes_,
            "G_allele_loci_per_s_not_pipelined": R * N * L * P / ((ms_sim + ms_all) * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "k_emit_rows<2>, one launch", "achieved": bytes_ / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_ / (k_ms * 1e-3) / 1e9 / peak}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pedigreesim_b200 as ps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libpsim has no CPU fallback)")
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workload()
    P, L, N, M = w["P"], w["L"], w["N"], w["M"]
    R = max(1, args.replicates)   # replicate populations a context holds and a step simulates + emits (side by side in the tables)
    fp = (R * N * P + 127) // 128 * 128
    dp = (R * N + 127) // 128 * 128

    # three contexts = three device-resident populations in rotation: while population k is emitted, population k+1 is
    # simulated, and the context that takes population k+1 finished its last emission a whole step ago (no wait)
    NCTX = 3
    ctxs, streams, tgts, keep = [], [], [], []
    for k in range(NCTX):
        ctx = ps.PsimContext(P, seed=SEED, replicate_first=rank * R, replicate_count=R, device=local, **w["model"])
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        ctx.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
        ctx.set_pedigree(w["p0"], w["p1"])
        ctx.set_map(w["off"], w["pos"])
        ctx.set_allele_codes(w["code"])
        d_founder = torch.empty((L, fp), dtype=torch.uint8, device="cuda")
        d_dose = torch.empty((L, dp), dtype=torch.uint8, device="cuda")
        tgt = ps.EmitTargets()
        tgt.founder, tgt.founder_pitch, tgt.founder_bytes = d_founder.data_ptr(), fp, 1
        tgt.dose, tgt.dose_pitch, tgt.dose_which = d_dose.data_ptr(), dp, ps.DOSE_MAX_ALLELE
        tgt.memory = ps.MEM_DEVICE
        tgt.flags = ps.EMIT_ASYNC
        ctxs.append(ctx); streams.append(stream); tgts.append(tgt); keep.append((d_founder, d_dose))
    ctx = ctxs[0]

    def barrier():
        if world > 1:
            dist.barrier()

    acc = {"sim": 0.0, "emit": 0.0, "emit_total": 0.0, "emit_launches": 0, "n": 0}

    def collect(c):   # device times of the context's last simulate + emit (both complete)
        c.wait()
        st = c.stats()
        acc["sim"] += st["ms_simulate"]
        acc["emit"] += st["ms_emit_kernel"]
        acc["emit_total"] += st["ms_emit_total"]
        acc["emit_launches"] += st["emit_kernel_launches"]
        acc["n"] += 1

    def replicate_id(k):
        return (rank + world * (k + 1)) * R   # fresh replicates every step: nothing is cached

    def run_steps(first, count, timed, targets=None):
        targets = targets or tgts
        # step k: population k is emitted (its simulation was enqueued one step earlier) while population k+1 is simulated
        ctxs[first % NCTX].reset(SEED, replicate_id(first))
        ctxs[first % NCTX].simulate_async()
        for k in range(first, first + count):
            a, b = k % NCTX, (k + 1) % NCTX
            if timed and k - (NCTX - 1) >= first:
                collect(ctxs[b])                       # its emission of step k - 2 finished long ago: no wait
            if args.order == "emit-first":
                ctxs[a].emit_into(targets[a], 0, N, 0, L)  # the one host wait of the step: population k's simulation; then K0 + K2 are enqueued
                ctxs[b].reset(SEED, replicate_id(k + 1))
                ctxs[b].simulate_async()
                continue
            ctxs[b].reset(SEED, replicate_id(k + 1))
            ctxs[b].simulate_async()
            ctxs[a].emit_into(targets[a], 0, N, 0, L)  # the one host wait of the step: population k's simulation; then K0 + K2 are enqueued
        for c in ctxs:
            c.wait()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    run_steps(0, args.warmup, False)
    if rank == 0:
        sampler.wait_for(1, 3.0)   # nvidia-smi is up and sampling every 100 ms
    l0 = sum(c.stats()["kernel_launches"] for c in ctxs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    s0 = sampler.mark()
    e0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(e0)
    run_steps(args.warmup, args.steps, True)
    for st_ in streams[1:]:
        streams[0].wait_stream(st_)
    e1.record(streams[0])
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(c.stats()["kernel_launches"] for c in ctxs) - l0
    clocks = None
    if rank == 0:
        # nvidia-smi samples every 100 ms and a step lasts half a millisecond: when the timed region was too short for
        # three samples, the very same steps go on (untimed) right behind it until the window holds three
        t_roll, k_roll = time.time(), args.warmup + args.steps
        while sampler.proc and sampler.mark() - s0 < 3 and time.time() - t_roll < 2.0:
            run_steps(k_roll, 20, False)
            k_roll += 20
        clocks = sampler.stop(s0)
        clocks["window"] = "the timed region (%.1f ms) and the identical untimed steps that follow it until nvidia-smi (-lms 100) has given three samples" % ms
    # the simulation of the population after the last one was enqueued too (it is part of the timed work: K steps emit K
    # populations and simulate K + 1); kernel times of the roofline kernel, alone on the device:
    solo = {"emit": 0.0, "launches": 0, "sim": 0.0, "total": 0.0}
    tgts[0].flags = 0
    for k in range(3):
        ctx.reset(SEED, replicate_id(100000 + k))
        ctx.simulate()
        ctx.emit_into(tgts[0], 0, N, 0, L)
        st = ctx.stats()
        solo["emit"] += st["ms_emit_kernel"]; solo["launches"] += st["emit_kernel_launches"]
        solo["sim"] += st["ms_simulate"]; solo["total"] += st["ms_emit_total"]
    tgts[0].flags = ps.EMIT_ASYNC

    # ---- the same pipelined step with PSIM_EMIT_PACKED4 device tables (two 4-bit cells per byte: half the HBM stream)
    pfp, pdp = ((R * N * P + 1) // 2 + 127) // 128 * 128, ((R * N + 1) // 2 + 127) // 128 * 128
    ptgts, pkeep = [], []
    for k in range(NCTX):
        pf = torch.empty((L, pfp), dtype=torch.uint8, device="cuda")
        pd = torch.empty((L, pdp), dtype=torch.uint8, device="cuda")
        t = ps.EmitTargets()
        t.founder, t.founder_pitch, t.founder_bytes = pf.data_ptr(), pfp, 1
        t.dose, t.dose_pitch, t.dose_which = pd.data_ptr(), pdp, ps.DOSE_MAX_ALLELE
        t.memory = ps.MEM_DEVICE
        t.flags = ps.EMIT_ASYNC | ps.EMIT_PACKED4
        ptgts.append(t); pkeep.append((pf, pd))
    run_steps(200000, 3, False, ptgts)
    barrier()
    torch.cuda.synchronize()
    e0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(e0)
    run_steps(200003, args.steps, False, ptgts)
    for st_ in streams[1:]:
        streams[0].wait_stream(st_)
    e1.record(streams[0])
    torch.cuda.synchronize()
    barrier()
    ms_packed = e0.elapsed_time(e1)
    del pkeep

    # ---- end to end through the C ABI with host buffers: one population per step, in a context of its own
    stream = streams[0]
    ctx_e = ps.PsimContext(P, seed=SEED, replicate_first=rank, device=local, **w["model"])
    ctx_e.set_stream(stream.cuda_stream)
    ctx_e.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
    e2e_steps = max(1, min(args.steps, 5))
    trace = bool(os.environ.get("PSIM_BENCH_TRACE"))

    def e2e_leg(packed):
        wf, wd = ((N * P + 1) // 2, (N + 1) // 2) if packed else (N * P, N)
        h_founder = torch.empty((L, wf), dtype=torch.uint8, pin_memory=True)
        h_dose = torch.empty((L, wd), dtype=torch.uint8, pin_memory=True)
        ht = ps.EmitTargets()
        ht.founder, ht.founder_pitch, ht.founder_bytes = h_founder.data_ptr(), wf, 1
        ht.dose, ht.dose_pitch, ht.dose_which = h_dose.data_ptr(), wd, ps.DOSE_MAX_ALLELE
        ht.memory = ps.MEM_HOST
        ht.flags = ps.EMIT_PACKED4 if packed else 0

        def e2e_step(k):
            ts = [time.perf_counter()]
            ctx_e.set_pedigree(w["p0"], w["p1"])  # drops the population and re-uploads the pedigree
            ts.append(time.perf_counter())
            ctx_e.reset(SEED, rank + world * (1000 + k))
            ctx_e.set_map(w["off"], w["pos"])
            ctx_e.set_allele_codes(w["code"])
            ts.append(time.perf_counter())
            ctx_e.simulate()
            ts.append(time.perf_counter())
            ctx_e.emit_into(ht, 0, N, 0, L)
            ts.append(time.perf_counter())
            if trace:
                print("e2e step %d: set_pedigree %.2f inputs %.2f simulate %.2f emit %.2f ms" % (
                    (k,) + tuple((b - a) * 1e3 for a, b in zip(ts, ts[1:]))), file=sys.stderr)

        for k in range(3):   # W >= 3 also here: a fresh box's first pinned D2H copies are slow
            e2e_step(-1 - k)
        barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for k in range(e2e_steps):
            e2e_step(k)
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1), int(L * wf + L * wd)

    ms_e2e, d2h = e2e_leg(True)
    ms_e2e_u8, d2h_u8 = e2e_leg(False)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_u8, ms_packed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_u8, ms_packed = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    sharded = None
    if world > 1 and args.sharded_lines > 0:   # every rank takes part; rank 0 reports
        try:
            sharded = sharded_block(ps, torch, dist, rank, world, local, args.sharded_lines)
        except Exception as e:   # (never let the extra block take the line down)
            sharded = {"error": str(e)}
    ctx_e.close()
    cells = R * N * L * P
    value = world * cells * args.steps / (ms * 1e-3) / 1e9
    e2e_value = world * N * L * P * e2e_steps / (ms_e2e * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    k2_ms = solo["emit"] / max(1, solo["launches"])
    algo_bytes = R * N * L * (P + 1)
    achieved = algo_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    h2d = int(w["p0"].nbytes * 2 + w["off"].nbytes + w["pos"].nbytes + w["code"].nbytes)
    if rank == 0:
        out = {
            "metric": "genotype calls/sec (G allele-loci/s)", "value": value, "unit": "G allele-loci/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "per_gpu": "%d replicate population(s) of the workload per rank and step, side by side in one context "
                                                                     "(replicate_count = %d) and in its tables; no collective" % (R, R),
                       "replicates_per_step": R,
                       "tables": "founder-allele u8 [L][R*N*P] + MAX-allele dose u8 [L][R*N]",
                       "pipeline": "three contexts per rank in rotation: population k+1 is simulated (psim_simulate_async) while population k is "
                                   "emitted (PSIM_EMIT_ASYNC); one host wait per step (K steps emit K populations and simulate K + 1)",
                       "l2": "%.1f GB of output per step exceeds the 126 MB L2; new replicates are simulated every step" % (R * N * L * (P + 1) / 1e9),
                       "e2e_tables": "PSIM_EMIT_PACKED4: founder-allele and dose tables with two 4-bit cells per byte (e2e.u8: one byte per cell)",
                       "positions": "fp64 Morgan"},
            "meioses_per_s": world * R * M * args.steps / (ms * 1e-3),
            "kernel_ms": {"simulate_per_step_overlapped": acc["sim"] / max(1, acc["n"]),
                          "emit_all_kernels_per_step_overlapped": acc["emit_total"] / max(1, acc["n"]),
                          "simulate_alone": solo["sim"] / 3, "emit_all_kernels_alone": solo["total"] / 3,
                          "emit_kernel_per_launch": k2_ms,
                          "emit_kernel_share_of_step": k2_ms / (ms / args.steps) if ms > 0 else None},
            "roofline": {"bound": "hbm", "kernel": "k_emit_rows<2> (row-image emission: shared-memory row + TMA bulk store per table row; one launch emits the "
                                                   "step's %d population(s); timed alone on the device, 3 launches after the pipelined steps; the band-plan helper "
                                                   "kernels are reported under kernel_ms)" % R,
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_launch": algo_bytes,
                         "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": "G allele-loci/s", "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "tables": "PSIM_EMIT_PACKED4",
                    "workload": "one population of the workload per step (host tables of one population)",
                    "u8": {"value": world * N * L * P * e2e_steps / (ms_e2e_u8 * 1e-3) / 1e9, "ms_per_step": ms_e2e_u8 / e2e_steps,
                           "d2h_bytes_per_step": d2h_u8}},
            "packed_tables": {"value": world * cells * args.steps / (ms_packed * 1e-3) / 1e9, "unit": "G allele-loci/s", "ms_per_step": ms_packed / args.steps,
                              "tables": "PSIM_EMIT_PACKED4 device tables (4 bits per cell): the same pipelined step, half the bytes per step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if sharded is not None:
            out["sharded"] = sharded
        if world == 1:
            try:   # the other batch size, not pipelined: one population per step when the line is about eight, and the reverse
                other = 1 if R > 1 else 8
                out["single_population" if other == 1 else "replicate_batch"] = replicate_batch(ps, torch, local, rank, w, other)
            except Exception as e:   # (never let the extra block take the line down)
                out["single_population" if R > 1 else "replicate_batch"] = {"error": str(e)}
            try:
                out["replicate_sweep"] = replicate_sweep(ps, torch, local, rank)
            except Exception as e:   # (never let the extra block take the line down)
                out["replicate_sweep"] = {"error": str(e)}
        if world == 1 and not args.no_cpu:
            v, t, ws = cpu_port_run(w, CPU_SAMPLE_POPSIZE, 1, 0)
            out["cpu_baseline"] = {"value": v, "unit": "G allele-loci/s", "cores": 1, "kind": "port",
                                   "sample": "same workload with POPSIZE=%d of 10000 offspring, all 24000 loci, "
                                             "once: %.1f s" % (CPU_SAMPLE_POPSIZE, t),
                                   "meioses_per_s": ws["M"] / t}
        prof = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(prof):
            try:
                tr = json.load(open(prof))
                if tr.get("algorithmic_bytes_per_launch") == algo_bytes:   # (a capture of this very launch geometry)
                    out["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
                    out["roofline"]["traffic_source"] = "not measured in this run: " + tr.get("source", "profiles/k2_traffic.json")
            except Exception:
                pass
        print(json.dumps(out))
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 200 (a step is half a millisecond); --impl reference: 2 (a step is seconds)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--replicates", type=int, default=8,
                    help="replicate populations of the workload a rank simulates and emits per step, side by side in one context "
                         "(replicate_count): a Monte Carlo sweep over the workload, the job replicate sharding is for. 1: one population per step")
    ap.add_argument("--order", default="simulate-first", choices=["simulate-first", "emit-first"],
                    help="what a pipelined step enqueues first: the next population's meioses or this population's emission")
    ap.add_argument("--sharded-lines", type=int, default=1000000,
                    help="N > 1: F2 lines of the one-pedigree-over-all-GPUs block (configs[3]: 1 000 000); 0 skips it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.steps is None:
        args.steps = 2 if args.impl == "reference" else 200
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
