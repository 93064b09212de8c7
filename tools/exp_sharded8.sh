#!/bin/bash
# 8 GPUs: psim_simulate_sharded on configs[3] (both partitions, both transfer kinds) and the bench line at N = 8
N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/run_sharded.py "${@:2}" --out gpurun_out/r2_sharded$N.jsonl 2>> gpurun_out/r2_sharded$N.err | tail -1 | cut -c1-420; }
run 29515 --lines 1000000 --generations 8 --partition block --impl library
run 29516 --lines 1000000 --generations 8 --partition lineage --impl library
echo "== ncclBroadcast instead of send / recv"
PSIM_TUNING_LIB=build/libpsim_tune.so PSIM_SHARD_BCAST=1 run 29518 --lines 1000000 --generations 8 --partition block --impl library
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err
tail -1 gpurun_out/r2_bench_n8.log | cut -c1-300
tail -3 gpurun_out/r2_sharded$N.err
