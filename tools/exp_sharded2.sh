#!/bin/bash
# N GPUs: the NCCL transport of psim_simulate_sharded (tests + timing)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -2; bash tools/exp_host2.sh; fi
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tools/run_sharded.py "${@:2}" --out gpurun_out/r2_sharded${N}b.jsonl 2>> gpurun_out/r2_sharded$N.err | tail -1 | cut -c1-420; }
run 29511 --lines 200000 --generations 8 --partition block --impl library
run 29512 --lines 200000 --generations 8 --partition lineage --impl library
run 29515 --lines 1000000 --generations 8 --partition block --impl library
run 29516 --lines 1000000 --generations 8 --partition lineage --impl library
tail -3 gpurun_out/r2_sharded$N.err
