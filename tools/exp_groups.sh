#!/bin/bash
# Row-image emission: one / two groups of compute threads (PSIM_ROWS_GROUPS) x one issuing warp per buffer or a single
# issuing warp with at most k rows in the bulk-copy engine (PSIM_ROWS_INFLIGHT); configs[1] through tools/prof_step.py
mkdir -p gpurun_out
{
for lib in build/libpsim_g1f0.so build/libpsim_g2f0.so build/libpsim_g1f1.so build/libpsim_g1f2.so build/libpsim_g2f1.so build/libpsim_g2f2.so build/libpsim_g2f3.so; do
  for opts in "" "--packed" "--no-codes"; do
    echo "== $lib $opts"
    PSIM_TUNING_LIB=$lib timeout 120 python tools/prof_step.py --steps 6 $opts 2>&1 | tail -2
  done
  PSIM_TUNING_LIB=$lib timeout 600 python -m pytest tests/test_gpu_emit.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
done
} > gpurun_out/r2_groups2.log 2>&1
cat gpurun_out/r2_groups2.log
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/r2_shard_tests.log 2>&1
tail -30 gpurun_out/r2_shard_tests.log
