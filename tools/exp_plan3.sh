#!/bin/bash
mkdir -p gpurun_out
{
for sh in -1 1; do
echo "== PSIM_PLAN_SHARED=$sh"
PSIM_PLAN_SHARED=$sh PSIM_TUNING_LIB=build/libpsim_tune.so timeout 300 python bench.py --steps 40 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f kernel_ms %s sweep %s' % (d['value'], d['ms_per_step'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), json.dumps({k: round(d['replicate_sweep'][k],3) for k in ('simulate_ms','emit_kernel_ms','emit_all_kernels_ms')})))"
done
echo "== shipped"
timeout 300 python bench.py --steps 40 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f kernel_ms %s sweep %s' % (d['value'], d['ms_per_step'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), json.dumps({k: round(d['replicate_sweep'][k],3) for k in ('simulate_ms','emit_kernel_ms','emit_all_kernels_ms')})))"
} > gpurun_out/r2_plan3.log 2>&1
cat gpurun_out/r2_plan3.log
