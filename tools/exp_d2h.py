"""PCIe D2H ceiling experiments: whole-table copy vs chunked copies, with / without NUMA binding."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

if len(sys.argv) > 1 and sys.argv[1] == "bind":
    bench.bind_to_gpu_numa_node(0)
print("affinity:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
n = 1200 * 1000 * 1000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def t(label, f, reps=4):
    for k in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("%-40s %.2f ms = %.1f GB/s" % (label, (t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
t("whole .pin_memory()", lambda: h.copy_(d, non_blocking=True))
t("whole pin_memory=True", lambda: h2.copy_(d, non_blocking=True))
def chunked(hh, cs):
    def f():
        for o in range(0, n, cs):
            hh[o:o + cs].copy_(d[o:o + cs], non_blocking=True)
    return f
for cs in (16 << 20, 64 << 20, 192 << 20):
    t("chunks of %d MiB" % (cs >> 20), chunked(h2, cs))
s2 = torch.cuda.Stream()
def two_streams():
    cs = 192 << 20
    for k, o in enumerate(range(0, n, cs)):
        with torch.cuda.stream(s2 if k & 1 else torch.cuda.current_stream()):
            h2[o:o + cs].copy_(d[o:o + cs], non_blocking=True)
t("192 MiB chunks alternating 2 streams", two_streams)
