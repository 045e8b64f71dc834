(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python tools/bench_single_frame.py 2>/dev/null | tail -1 | cut -c1-220
python tools/latency_breakdown.py 2>/dev/null | tail -5
