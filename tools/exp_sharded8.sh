# one large pedigree sharded per generation over N GPUs with an NCCL all-gather of the new HaploStructs
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/run_sharded.py --lines 200000 --generations 8 2>gpurun_out/sharded_err_$n.log | tail -1 >> gpurun_out/sharded_r1s3.jsonl
done
