#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err
tail -1 gpurun_out/r2_bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['packed_tables']['value'], d['roofline']['frac'], json.dumps({k: (d['sharded'][k]['seconds'], d['sharded'][k]['speedup_vs_one_gpu'], d['sharded'][k]['sample_bit_identical_to_one_gpu']) for k in ('block','lineage')}))"
tail -2 gpurun_out/r2_bench_n2.err
