#!/bin/bash
# configs[2..4] at full size (tools/scale_configs.py) and compute-sanitizer memcheck over the small GPU tests
mkdir -p gpurun_out
timeout 900 python tools/scale_configs.py 3 4 5 > gpurun_out/r2_scale.jsonl 2> gpurun_out/r2_scale.err
cut -c1-700 gpurun_out/r2_scale.jsonl
tail -3 gpurun_out/r2_scale.err
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_shard.py tests/test_gpu_emit.py -x -q -m gpu -k "not full_size and not nccl" > gpurun_out/r2_sanitizer.log 2>&1
echo "sanitizer rc $?"
tail -6 gpurun_out/r2_sanitizer.log
