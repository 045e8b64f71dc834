# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is tools/collect_profiles.sh <tag>
TAG=r2b
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 200 python tools/tc_debug.py 2>&1 | tail -25) > gpurun_out/${TAG}_tc_vs_ffma.log 2>&1
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
(DCU_SEG=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_noseg.json 2>gpurun_out/${TAG}_bench_noseg.err
(DCU_SLICE_MINOR=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_slicemajor.json 2>gpurun_out/${TAG}_bench_slicemajor.err
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench2.json 2>gpurun_out/${TAG}_bench2.err
(timeout 900 python tools/parity_report.py --frames 2048 --impls tcgen05 --out gpurun_out/${TAG}_parity_2048.json 2>&1 | tail -30) > gpurun_out/${TAG}_parity_2048.log 2>&1
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_tc_vs_ffma.log
for f in bench bench_noseg bench_slicemajor bench2; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$f.json")); print("$f", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "conv ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["clocks"])
except Exception as e: print("$f", "ERR", e); print(open("gpurun_out/${TAG}_$f.err").read()[-2000:])
PY
done
tail -12 gpurun_out/${TAG}_parity_2048.log
