# weak scaling over replicate populations: bench.py at N = 8, 4, 2, 1 on one box
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu 2>gpurun_out/scale_s3_err_$n.log | tail -1 >> gpurun_out/scale_s3.jsonl
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/scale_s3.jsonl
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
