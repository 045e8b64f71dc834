# scratch command file for `gpurun -- 'bash tools/_run.sh'`
TAG=r2m
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for i in 1 2; do (timeout 200 python tools/bench_single_frame.py 2>&1 | tail -1 | cut -c90-220); done
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_list.log 2>&1
python - <<PY
import csv
rows=[l for l in open("gpurun_out/${TAG}_launches.csv") if not l.startswith("==")]
tot=0
for r in csv.DictReader(rows):
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]; tot += v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
print("step kernel us (ncu):", round(tot,1))
PY
for i in a b; do
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1) > gpurun_out/${TAG}_bench_$i.json 2>gpurun_out/${TAG}_bench_$i.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$i.json")); print("$i", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "py", round(d["e2e_python"]["value"]), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print("$i", "ERR", e); print(open("gpurun_out/${TAG}_bench_$i.err").read()[-1500:])
PY
done
