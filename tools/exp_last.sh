#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_text.py tests/test_gpu_host.py tests/test_gpu_configs.py tests/test_gpu_stats.py -x -q -m gpu 2>&1 | tail -1
PSIM_BENCH_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_trace_bench.log 2> gpurun_out/r2_trace_bench.err
grep "e2e step" gpurun_out/r2_trace_bench.err | sed -n 4,8p
tail -1 gpurun_out/r2_trace_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:400])"
