#!/bin/bash
mkdir -p gpurun_out
{
for lib in build/libpsim_pcall.so build/libpsim_logcall.so build/libpsim_trace.so; do
  echo "== $lib"; PSIM_TUNING_LIB=$lib timeout 120 python tools/prof_step.py --steps 8 2>&1 | tail -3
  PSIM_TUNING_LIB=$lib timeout 300 python tools/prof_sweep.py 2>&1 | tail -1 | cut -c100-260
done
timeout 900 python -m pytest tests/test_gpu_meiosis.py tests/test_gpu_stats.py tests/test_gpu_shard.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/r2_k1b.log 2>&1
cat gpurun_out/r2_k1b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_emit_rows$ -s 3 -c 1 -f -o gpurun_out/r2_prof_rows python tools/prof_step.py --steps 4 > gpurun_out/r2_prof_rows_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_emit_rows_packed$ -s 3 -c 1 -f -o gpurun_out/r2_prof_rows_packed python tools/prof_step.py --steps 4 --packed > gpurun_out/r2_prof_rows_packed_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_rows_ncu.log
