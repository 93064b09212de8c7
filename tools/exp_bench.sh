# round-end style run: tests, smoke, both bench arms, then the ncu launch list of the bench command and one full capture of the roofline kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_tests_s3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke_s3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke_s3.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1s3.log 2>&1
python bench.py > gpurun_out/bench_plain_r1s3.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1s3.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_r1s3.log 2>&1
python tools/prof_step.py --steps 3 > gpurun_out/prof_plain_s3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_emit_rows -c 2 -o gpurun_out/prof_emit_r1s3 -f python tools/prof_step.py --steps 3 > gpurun_out/prof_ncu_s3.log 2>&1
