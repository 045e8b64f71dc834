(timeout 600 python -m pytest tests/test_gpu_parity_oracle.py -m gpu -x -q -k awkward -s 2>&1 | tail -15)
