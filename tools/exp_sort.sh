#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r2_sort_tests.log 2>&1
tail -3 gpurun_out/r2_sort_tests.log
{
PSIM_ROWS_TRACE=1 PSIM_TUNING_LIB=build/libpsim_tune.so timeout 300 python tools/prof_sweep.py 2>&1 | tail -3
for lib in build/libpsim_tune.so build/libpsim_pcall.so; do
for opts in "" "--packed"; do echo "== $lib $opts"; PSIM_TUNING_LIB=$lib timeout 120 python tools/prof_step.py --steps 6 $opts 2>&1 | tail -2; done
echo "== sweep $lib"; PSIM_TUNING_LIB=$lib timeout 300 python tools/prof_sweep.py 2>&1 | tail -1 | cut -c1-330
done
PSIM_TUNING_LIB=build/libpsim_pcall.so timeout 600 python -m pytest tests/test_gpu_meiosis.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/r2_sort.log 2>&1
cat gpurun_out/r2_sort.log
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sweep_launches3.csv python tools/prof_sweep.py > gpurun_out/r2_sweep_ncu3.log 2>&1
