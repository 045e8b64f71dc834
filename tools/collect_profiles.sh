#!/bin/bash
# Run ON THE GPU BOX (under gpurun): tests, bench (both arms), parity report, ncu launch list and ncu full captures of ONE step of the
# bench workload (batch 256, engine defaults) -- every kernel instantiation the step runs.
# Usage: bash tools/collect_profiles.sh <tag> [parity_frames]      -> writes gpurun_out/<tag>_*
TAG=${1:-rX}
PF=${2:-4096}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${TAG}_smoke.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
(timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1) > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
kill $SMI
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/${TAG}_bench_reference.json 2>&1
(DCU_SEG=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1) > gpurun_out/${TAG}_bench_strict.json 2>gpurun_out/${TAG}_bench_strict.err
# every launch of one step with its device time and DRAM bytes (shares of the step)
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_list.log 2>&1
# full metric set for every launch of the same step (report stays on the box; the raw page comes back as CSV)
(timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_step_full python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_full.log 2>&1
(ncu -i /tmp/${TAG}_step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv) 2>> gpurun_out/${TAG}_ncu_full.log
# the top kernel (conv1b: first conv_tc2 launch of the step) with source correlation
(timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -c 1 -f -o gpurun_out/${TAG}_conv1b python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_conv1b.log 2>&1
(timeout 200 python tools/layer_table.py --batch 256 --json gpurun_out/${TAG}_layer_table.json 2>&1 | tail -40) > gpurun_out/${TAG}_layer_table.log 2>&1
(timeout 200 python tools/tc_debug.py 2>&1 | tail -25) > gpurun_out/${TAG}_tc_vs_ffma.log 2>&1 || echo "tc_debug reported MISMATCH" >> gpurun_out/${TAG}_tc_vs_ffma.log
(timeout 300 python tools/mma_probe.py --out gpurun_out/${TAG}_mma_probe.json 2>&1 | tail -80) > gpurun_out/${TAG}_mma_probe.log 2>&1
(timeout 1500 python tools/parity_report.py --frames $PF --impls tcgen05 --out gpurun_out/${TAG}_parity.json 2>&1 | tail -40) > gpurun_out/${TAG}_parity.log 2>&1
(DCU_SEG=1 timeout 1500 python tools/parity_report.py --frames 1024 --impls tcgen05 --out gpurun_out/${TAG}_parity_strict.json 2>&1 | tail -12) > gpurun_out/${TAG}_parity_strict.log 2>&1
(timeout 600 python tools/parity_report.py --frames 256 --impls tcgen05,ffma --size 640x480 --out gpurun_out/${TAG}_parity_640.json 2>&1 | tail -12) > gpurun_out/${TAG}_parity_640.log 2>&1
bash tools/sanitize.sh ${TAG} 2>&1 | grep -E "==|SUMMARY"
tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_reference.json; tail -6 gpurun_out/${TAG}_parity.log; tail -3 gpurun_out/${TAG}_parity_strict.log; tail -3 gpurun_out/${TAG}_parity_640.log; ls -la gpurun_out/${TAG}_*
