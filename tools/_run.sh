export DCU_FUSE_FIRST=1
(timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q -k "fused_first" 2>&1 | tail -3)
b() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1), round(d['roofline']['issued_tflops'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; }
b fused16
DCU_FUSE_FIRST=0 b unfused
b fused16
python tools/layer_table.py --batch 256 > gpurun_out/s17_layers.log 2>&1; head -2 gpurun_out/s17_layers.log | cut -c1-200; tail -1 gpurun_out/s17_layers.log
