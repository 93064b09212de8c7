#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err
tail -1 gpurun_out/r2_bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['packed_tables'], json.dumps(d['sharded'])[:900])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
tail -3 gpurun_out/r2_bench_n2.err
