#!/bin/bash
# 8 GPUs: psim_simulate_sharded on configs[3], both partitions
N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/run_sharded.py "${@:2}" --out gpurun_out/r2_sharded${N}b.jsonl 2>> gpurun_out/r2_sharded$N.err | tail -1 | cut -c1-420; }
run 29515 --lines 1000000 --generations 8 --partition block --impl library
run 29516 --lines 1000000 --generations 8 --partition lineage --impl library
tail -3 gpurun_out/r2_sharded$N.err
