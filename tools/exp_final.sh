#!/bin/bash
# Final validation of the shipped library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_tests.log 2>&1
tail -2 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err || exit 1
tail -1 gpurun_out/r2_final_bench.log | cut -c1-260
