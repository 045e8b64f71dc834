# scratch command file for `gpurun -- 'bash tools/_run.sh'`
TAG=r2i
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_list.log 2>&1
for i in a b c; do
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1) > gpurun_out/${TAG}_bench_$i.json 2>gpurun_out/${TAG}_bench_$i.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$i.json")); print("$i", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "py", round(d["e2e_python"]["value"]), "conv ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "issued", round(d["roofline"]["issued_frac"],3), d["clocks"]["sm_mhz"])
except Exception as e: print("$i", "ERR", e); print(open("gpurun_out/${TAG}_bench_$i.err").read()[-1500:])
PY
done
