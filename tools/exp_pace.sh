#!/bin/bash
# (a) k_emit_rows with the SMs' row hand-over paced to an aggregate rate (PSIM_ROWS_PACE, GB/s); (b) k_meiosis with
# different units per block / block sizes; (c) the meiosis parity tests on the shipped library
mkdir -p gpurun_out
{
for pace in 0 5600 6000 6300 6600 6900 7300 8000; do
  for opts in "" "--packed"; do
    echo "== pace $pace $opts"
    PSIM_ROWS_PACE=$pace PSIM_TUNING_LIB=build/libpsim_tune.so timeout 120 python tools/prof_step.py --steps 6 $opts 2>&1 | tail -2
  done
done
for upb in -1 256 224 192 160 128; do
  echo "== threads 256 upb $upb"
  PSIM_MEI_UPB=$upb PSIM_TUNING_LIB=build/libpsim_tune.so timeout 120 python tools/prof_step.py --steps 6 2>&1 | tail -2
done
for upb in -1 128 96; do
  echo "== threads 128 upb $upb"
  PSIM_MEI_UPB=$upb PSIM_TUNING_LIB=build/libpsim_tune128.so timeout 120 python tools/prof_step.py --steps 6 2>&1 | tail -2
done
} > gpurun_out/r2_pace.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_meiosis.py tests/test_gpu_stats.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/r2_pace_tests.log 2>&1
tail -5 gpurun_out/r2_pace_tests.log
cat gpurun_out/r2_pace.log
