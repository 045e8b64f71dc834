(timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q -k "argmax_heads" 2>&1 | tail -3)
