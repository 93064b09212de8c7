#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.log 2> gpurun_out/r2_final_bench.err || exit 1
tail -1 gpurun_out/r2_final_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['packed_tables']['value'], json.dumps(d['replicate_batch'])[:700], json.dumps(d['cpu_baseline'])[:200])"
tail -3 gpurun_out/r2_final_bench.err
