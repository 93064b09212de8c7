timeout 600 python -m pytest tests/test_gpu_emit.py -m gpu -x -q -k "overflowing" 2>&1 | tail -30
