# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is tools/collect_profiles.sh <tag>
TAG=r2d
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log 2>&1
run_bench() {  # name, env...
  name=$1; shift
  (env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_$name.json 2>gpurun_out/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$name.json")); print("$name", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "conv ms", round(d["roofline"]["kernel_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["clocks"])
except Exception as e: print("$name", "ERR", e); print(open("gpurun_out/${TAG}_bench_$name.err").read()[-1500:])
PY
}
run_par() { name=$1; shift
  (env "$@" timeout 600 python tools/parity_report.py --frames 256 --impls tcgen05 --out gpurun_out/${TAG}_parity_$name.json 2>&1 | grep "max_abs_dloc" | cut -c1-200) 2>&1 | sed "s/^/$name /"
}
run_bench seg_2_1 A=1
run_bench seg_2_1_cg1 DCU_SEG_CG=1
run_bench seg_2_2 DCU_SEG_CHUNKS128=2
run_bench seg_4_1 DCU_SEG_CHUNKS64=4
run_bench seg_1_1_first2 DCU_SEG_CHUNKS64=1 DCU_SEG_FIRST=2
run_bench noseg DCU_SEG=0
run_bench seg128only DCU_SEG=2
run_par seg_2_1 A=1
run_par seg_2_2 DCU_SEG_CHUNKS128=2
run_par seg_4_1 DCU_SEG_CHUNKS64=4
run_par noseg DCU_SEG=0
tail -8 gpurun_out/${TAG}_pytest.log
