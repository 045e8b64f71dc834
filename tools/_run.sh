DCU_NT64=2 python tools/layer_table.py --batch 256 --json gpurun_out/s9_layers.json > gpurun_out/s9_layers.log 2>&1; grep "conv3x3_tc" gpurun_out/s9_layers.log | cut -c1-200; tail -1 gpurun_out/s9_layers.log
b() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1), round(d['roofline']['issued_tflops'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; }
b base
DCU_NT64=2 b nt64all
b base
DCU_NT64=2 b nt64all
