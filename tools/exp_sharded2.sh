n=${1:-2}
for part in ${2:-lineage block}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/run_sharded.py --lines 200000 --generations 8 --partition $part 2>gpurun_out/sharded_err_$n.log | tail -1 >> gpurun_out/sharded_r1s3.jsonl
done
