#!/bin/bash
mkdir -p gpurun_out
{
for lib in build/libpsim_tune.so build/libpsim_near4.so build/libpsim_near8.so build/libpsim_near12.so; do
  echo "== $lib"; PSIM_TUNING_LIB=$lib timeout 120 python tools/prof_step.py --steps 8 2>&1 | tail -3 | cut -c1-80
  PSIM_TUNING_LIB=$lib timeout 300 python tools/prof_sweep.py 2>&1 | tail -1 | cut -c100-200
done
timeout 900 python -m pytest tests/test_gpu_meiosis.py tests/test_gpu_stats.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/r2_near.log 2>&1
cat gpurun_out/r2_near.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_near_bench.log 2> gpurun_out/r2_near_bench.err
tail -1 gpurun_out/r2_near_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['kernel_ms'], d['packed_tables'], d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/r2_near_bench.err
