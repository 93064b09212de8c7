#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py --steps 4 --replicates 8 > gpurun_out/r2_prof_step8.log 2>&1 || exit 1
tail -2 gpurun_out/r2_prof_step8.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_emit_rows$ -s 3 -c 1 -f -o gpurun_out/r2_prof_rows8 python tools/prof_step.py --steps 4 --replicates 8 > gpurun_out/r2_prof_rows8_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_prof_bench_ncu.log 2>&1
tail -1 gpurun_out/r2_prof_rows8_ncu.log
