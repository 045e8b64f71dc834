python tools/bench_configs.py --iters 5 --out gpurun_out/r1h_configs.json > gpurun_out/r1h_configs.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/r1h_configs.json'))
for k,v in d.items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items()})"
