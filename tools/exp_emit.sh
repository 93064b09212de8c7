timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; timeout 60 env "$@" python tools/prof_step.py --steps 6 2>&1 | tail -2; }
run X=1
run PSIM_SEG_MAX=20032 PSIM_ROWS_BLOCKS_PER_SM=2
run PSIM_SEG_MAX=20032 PSIM_ROWS_BLOCKS_PER_SM=2 PSIM_ROWS_THREADS=256
run PSIM_SEG_MAX=20032 PSIM_ROWS_BLOCKS_PER_SM=2 PSIM_BAND_ROWS=81
run PSIM_SEG_MAX=20032 PSIM_ROWS_BLOCKS_PER_SM=2 PSIM_BAND_ROWS=41
run PSIM_SEG_MAX=13376 PSIM_ROWS_BLOCKS_PER_SM=3 PSIM_ROWS_THREADS=320 PSIM_ROWS_CAP=4096
run PSIM_SEG_MAX=10048 PSIM_ROWS_BLOCKS_PER_SM=4 PSIM_ROWS_THREADS=256 PSIM_ROWS_CAP=4096
run PSIM_SEG_MAX=10048 PSIM_ROWS_BLOCKS_PER_SM=4 PSIM_ROWS_THREADS=256 PSIM_ROWS_CAP=4096 PSIM_BAND_ROWS=81
run PSIM_BAND_ROWS=81
run PSIM_BAND_ROWS=54
python - <<'PY'
import torch, time
n = 1200 * 1000 * 1000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for k in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("D2H 1.2 GB pinned: %.2f ms = %.1f GB/s" % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
PY
