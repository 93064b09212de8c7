#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_band_plan$ -s 2 -c 1 -f -o gpurun_out/r2_prof_plan python tools/prof_sweep.py > gpurun_out/r2_prof_plan_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_index_runs$ -s 2 -c 1 -f -o gpurun_out/r2_prof_index python tools/prof_sweep.py > gpurun_out/r2_prof_index_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_plan_ncu.log
