timeout 600 python -m pytest tests/test_gpu_stats.py -m gpu -x -q -k "device" 2>&1 | tail -30
