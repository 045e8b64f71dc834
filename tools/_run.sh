(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
python tools/layer_table.py --batch 256 --json gpurun_out/s6_layers.json > gpurun_out/s6_layers.log 2>&1; tail -22 gpurun_out/s6_layers.log | cut -c1-210
b() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1), round(d['roofline']['issued_tflops'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; }
b base
DCU_WRES=0 b nowres
b base
DCU_WRES=0 b nowres
