#!/bin/bash
mkdir -p gpurun_out
{
for lib in pedigreesim_b200/libpsim.so build/libpsim_B.so build/libpsim_C.so; do
  echo "== $lib (8 populations per step)"
  PSIM_TUNING_LIB=$lib timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f packed %.0f kernel_ms %s roofline %.3f' % (d['value'], d['ms_per_step'], d['packed_tables']['value'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), d['roofline']['frac']))"
done
} > gpurun_out/r2_reps4.log 2>&1
cat gpurun_out/r2_reps4.log
