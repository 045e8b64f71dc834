(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python tools/bench_single_frame.py 2>/dev/null | tail -1 | cut -c1-220
b() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['decode_gather'])"; }
b argheads
DCU_ARG_HEADS=0 b logits
b argheads
