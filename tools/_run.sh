(timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -5)
python tools/bench_single_frame.py 2>/dev/null | tail -1
DCU_GRAPH=0 python tools/bench_single_frame.py 2>/dev/null | tail -1
