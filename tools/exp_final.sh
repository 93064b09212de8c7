#!/bin/bash
# Final validation of the shipped library + the launch lists of record
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_tests.log 2>&1
tail -3 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err || exit 1
tail -1 gpurun_out/r2_final_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_benchref.log 2>&1
tail -1 gpurun_out/r2_final_benchref.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_prof_bench_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sweep_launches6.csv python tools/prof_sweep.py > gpurun_out/r2_sweep_ncu6.log 2>&1
bash tools/exp_host2.sh > /dev/null 2>&1
tail -4 gpurun_out/r2_host_program.txt | cut -c1-300
