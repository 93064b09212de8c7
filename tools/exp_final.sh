#!/bin/bash
# Final validation of the shipped library + the profiles of record
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_tests.log 2>&1
tail -3 gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err || exit 1
tail -1 gpurun_out/r2_final_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_benchref.log 2>&1
tail -1 gpurun_out/r2_final_benchref.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_prof_bench_ncu.log 2>&1
timeout 300 python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_step.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_emit_rows$ -s 3 -c 1 -f -o gpurun_out/r2_prof_rows python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_rows_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_meiosis -s 3 -c 1 -f -o gpurun_out/r2_prof_meiosis3 python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_meiosis3_ncu.log 2>&1
timeout 300 python tools/prof_step.py --steps 4 --packed > gpurun_out/r2_prof_step_packed.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_emit_rows_packed$ -s 3 -c 1 -f -o gpurun_out/r2_prof_rows_packed python tools/prof_step.py --steps 4 --packed > gpurun_out/r2_prof_rows_packed_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sweep_launches4.csv python tools/prof_sweep.py > gpurun_out/r2_sweep_ncu4.log 2>&1
tail -2 gpurun_out/r2_prof_step.log
