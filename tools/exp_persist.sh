#!/bin/bash
# k_meiosis as a grid of at most k blocks per SM (PSIM_MEI_PERSIST), so that an emission kernel always finds room next to it
mkdir -p gpurun_out
{
for k in 0 3 2; do
for R in 8 1; do
  echo "== PSIM_MEI_PERSIST=$k --replicates $R"
  PSIM_MEI_PERSIST=$k PSIM_TUNING_LIB=build/libpsim_tune.so timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --replicates $R 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f packed %.0f kernel_ms %s roofline %.3f' % (d['value'], d['ms_per_step'], d['packed_tables']['value'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), d['roofline']['frac']))"
done
done
PSIM_MEI_PERSIST=3 PSIM_TUNING_LIB=build/libpsim_tune.so timeout 600 python -m pytest tests/test_gpu_meiosis.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/r2_persist.log 2>&1
cat gpurun_out/r2_persist.log
