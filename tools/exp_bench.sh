one() { timeout 200 env "$@" python bench.py --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'], d['ms_per_step'])"; }
echo rows; one X=1; one X=1
echo rows-trace; one PSIM_BENCH_TRACE=1
echo norows; one PSIM_NO_ROWS=1; one PSIM_NO_ROWS=1
