# scratch command file for `gpurun -- 'bash tools/_run.sh'`
TAG=r2q
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
run_bench() {  # name, env...
  name=$1; shift
  (env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1) > gpurun_out/${TAG}_bench_$name.json 2>gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$name.json")); print("$name", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "py", round(d["e2e_python"]["value"]), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print("$name", "ERR", e); print(open("gpurun_out/${TAG}_bench_$name.err").read()[-1500:])
PY
}
run_bench devcount_a A=1
run_bench sync_a DCU_DEVICE_COUNT=0
run_bench devcount_b A=1
run_bench sync_b DCU_DEVICE_COUNT=0
run_bench devcount_c A=1
run_bench sync_c DCU_DEVICE_COUNT=0
