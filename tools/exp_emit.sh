timeout 300 python -m pytest tests/test_gpu_emit.py tests/test_gpu_meiosis.py -m gpu -x -q 2>&1 | tail -15
run() { echo "== $*"; timeout 60 env "$@" python tools/prof_step.py --steps 4 2>&1 | tail -2; }
run X=1
run PSIM_NO_ROWS=1
run PSIM_ROWS_BLOCKS_PER_SM=1
