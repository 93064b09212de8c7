timeout 600 python -m pytest tests/test_gpu_meiosis.py tests/test_gpu_stats.py -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; timeout 60 env "$@" python tools/prof_step.py --steps 5 2>&1 | tail -2; }
run X=1
