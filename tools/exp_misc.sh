#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -3
bash tools/exp_host2.sh
