#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_configs.py tests/test_gpu_text.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -3
{
timeout 300 python tools/prof_sweep.py 2>&1 | tail -1 | cut -c100-330
for opts in "" "--packed" "--no-codes"; do echo "== $opts"; timeout 120 python tools/prof_step.py --steps 6 $opts 2>&1 | tail -2; done
timeout 900 python tools/scale_configs.py 3 4 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:28], d['simulate_device_ms'], d['meioses_per_s'], d['emit_kernel_ms'], d['emit_all_kernels_ms'], d['rows_equal_tiled_kernel'])"
} > gpurun_out/r2_plan2.log 2>&1
cat gpurun_out/r2_plan2.log
