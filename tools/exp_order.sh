#!/bin/bash
mkdir -p gpurun_out
{
for order in simulate-first emit-first; do
  echo "== --order $order (8 populations per step)"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --order $order 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f packed %.0f kernel_ms %s' % (d['value'], d['ms_per_step'], d['packed_tables']['value'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()})))"
done
} > gpurun_out/r2_order.log 2>&1
cat gpurun_out/r2_order.log
