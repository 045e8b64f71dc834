(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
b() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])"; }
b chunked
DCU_CHUNKED_H2D=0 b onecopy
b chunked
DCU_CHUNKED_H2D=0 b onecopy
