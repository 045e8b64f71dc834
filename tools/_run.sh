(timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q -k "fused_first" 2>&1 | tail -8)
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
