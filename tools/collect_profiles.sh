#!/bin/bash
# Run ON THE GPU BOX (under gpurun): tests, bench, ncu launch list, ncu full captures, parity report.
# Usage: bash tools/collect_profiles.sh <tag>      -> writes gpurun_out/<tag>_*
TAG=${1:-rX}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${TAG}_smoke.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
(timeout 400 python bench.py --steps 30 --warmup 5 2>&1 | tail -1) > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
kill $SMI
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/${TAG}_bench_reference.json 2>&1
(timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_list.log 2>&1
(timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -c 8 -o gpurun_out/${TAG}_conv_tc python tools/profile_step.py --batch 32 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_conv.log 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"decode_gather|conv_first|heads_1x1" -c 3 -o gpurun_out/${TAG}_small python tools/profile_step.py --batch 256 2>&1 | tail -1) > gpurun_out/${TAG}_ncu_small.log 2>&1
(timeout 200 python tools/tc_debug.py 2>&1 | tail -20) > gpurun_out/${TAG}_tc_vs_ffma.log 2>&1
(timeout 600 python tools/parity_report.py --frames 256 --out gpurun_out/${TAG}_parity.json 2>&1 | tail -4) > gpurun_out/${TAG}_parity.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_reference.json; cat gpurun_out/${TAG}_parity.log
