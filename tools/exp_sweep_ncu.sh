#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_sweep.py > gpurun_out/r2_sweep_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sweep_launches.csv python tools/prof_sweep.py > gpurun_out/r2_sweep_ncu.log 2>&1
tail -2 gpurun_out/r2_sweep_plain.log
