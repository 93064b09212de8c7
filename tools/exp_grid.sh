#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python bench.py --steps 40 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f e2e %.1f kernel_ms %s sweep %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), json.dumps({k: round(d['replicate_sweep'][k],3) for k in ('simulate_ms','emit_kernel_ms','emit_all_kernels_ms')})))"
} > gpurun_out/r2_grid2.log 2>&1
cat gpurun_out/r2_grid2.log
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sweep_launches5.csv python tools/prof_sweep.py > gpurun_out/r2_sweep_ncu5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_step_launches.csv python tools/prof_step.py --steps 3 > /dev/null 2>&1
