#!/bin/bash
mkdir -p gpurun_out
for N in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err
tail -1 gpurun_out/r2_bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['packed_tables']['value'], json.dumps({k: (d['sharded'][k]['seconds'], d['sharded'][k]['speedup_vs_one_gpu'], d['sharded'][k]['sample_bit_identical_to_one_gpu']) for k in ('block','lineage')}))"
tail -2 gpurun_out/r2_bench_n$N.err
done
