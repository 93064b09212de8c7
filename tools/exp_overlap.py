"""Experiment: how much of a step overlaps when independent replicate populations are driven through
several libpsim contexts at once (one host thread and one CUDA stream per context; ctypes releases
the GIL). Prints steps/s and G allele-loci/s for 1, 2, 3 contexts on bench.py's configs[1] workload."""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pedigreesim_b200 as ps  # noqa: E402


def main():
    w = bench.workload()
    P, L, N = w["P"], w["L"], w["N"]
    fp = (N * P + 127) // 128 * 128
    dp = (N + 127) // 128 * 128
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    for n_ctx in (1, 2, 3):
        ctxs = []
        for k in range(n_ctx):
            ctx = ps.PsimContext(P, seed=bench.SEED, replicate_first=k, **w["model"])
            stream = torch.cuda.Stream()
            ctx.set_stream(stream.cuda_stream)
            ctx.set_chromosomes(w["lengths"], w["centro"], w["pref"], w["fq"])
            ctx.set_pedigree(w["p0"], w["p1"])
            ctx.set_map(w["off"], w["pos"])
            ctx.set_allele_codes(w["code"])
            f = torch.empty((L, fp), dtype=torch.uint8, device="cuda")
            d = torch.empty((L, dp), dtype=torch.uint8, device="cuda")
            t = ps.EmitTargets()
            t.founder, t.founder_pitch, t.founder_bytes = f.data_ptr(), fp, 1
            t.dose, t.dose_pitch, t.dose_which = d.data_ptr(), dp, ps.DOSE_MAX_ALLELE
            t.memory = ps.MEM_DEVICE
            ctxs.append((ctx, t, f, d, stream))

        def work(k, n):
            ctx, t = ctxs[k][0], ctxs[k][1]
            for s in range(n):
                ctx.reset(bench.SEED, 100 + k + n_ctx * s)
                ctx.simulate()
                ctx.emit_into(t, 0, N, 0, L)

        for k in range(n_ctx):
            work(k, 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(k, steps)) for k in range(n_ctx)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        total = n_ctx * steps
        print("contexts %d: %.3f ms per step, %.0f G allele-loci/s" % (n_ctx, dt / total * 1e3, total * N * L * P / dt / 1e9), flush=True)
        for c in ctxs:
            c[0].close()


if __name__ == "__main__":
    main()
