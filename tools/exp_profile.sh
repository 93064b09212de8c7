#!/bin/bash
# Round-2 profiles: plain bench, its ncu launch list, one ncu --set full capture each of k_emit_rows, k_emit_rows_packed
# and k_meiosis (every capture only after the same program has exited 0 without ncu)
mkdir -p gpurun_out
set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_prof_bench.log 2> gpurun_out/r2_prof_bench.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_prof_bench_ncu.log 2>&1
timeout 300 python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_step.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit_rows -s 3 -c 1 -f -o gpurun_out/r2_prof_rows python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_rows_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_meiosis -s 3 -c 1 -f -o gpurun_out/r2_prof_meiosis3 python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_meiosis3_ncu.log 2>&1
timeout 300 python tools/prof_step.py --steps 4 --packed > gpurun_out/r2_prof_step_packed.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit_rows_packed -s 3 -c 1 -f -o gpurun_out/r2_prof_rows_packed python tools/prof_step.py --steps 4 --packed > gpurun_out/r2_prof_rows_packed_ncu.log 2>&1
tail -1 gpurun_out/r2_prof_bench.log | cut -c1-400
