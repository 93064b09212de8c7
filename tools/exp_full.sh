#!/bin/bash
# Full GPU validation of the shipped library: every -m gpu test, smoke, the bench (both arms), the replicate sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_full_tests.log 2>&1
tail -3 gpurun_out/r2_full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_full_smoke.log 2>&1
tail -2 gpurun_out/r2_full_smoke.log
timeout 600 python bench.py > gpurun_out/r2_full_bench.log 2> gpurun_out/r2_full_bench.err
tail -1 gpurun_out/r2_full_bench.log | cut -c1-600
PSIM_ROWS_TRACE=1 timeout 300 python tools/prof_sweep.py 2>&1 | tail -2 | cut -c1-400
