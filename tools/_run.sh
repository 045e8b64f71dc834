(timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s4_pytest_tc.log 2>&1; tail -12 gpurun_out/s4_pytest_tc.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
python tools/layer_table.py --batch 256 --json gpurun_out/s4_layers.json > gpurun_out/s4_layers.log 2>&1; tail -22 gpurun_out/s4_layers.log | cut -c1-210
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/s4_bench.json; python -c "import json; d=json.load(open('gpurun_out/s4_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"
DCU_FLAT=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('noflat', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"
(timeout 600 python tools/parity_report.py --frames 256 --out gpurun_out/s4_parity.json 2>&1 | tail -4) > gpurun_out/s4_parity.log 2>&1; cat gpurun_out/s4_parity.log
