# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is tools/collect_profiles.sh <tag>
TAG=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
(timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 300 python tools/mma_probe.py --out gpurun_out/${TAG}_mma_probe.json 2>&1 | tail -80) > gpurun_out/${TAG}_mma_probe.log 2>&1
(timeout 200 python tools/tc_debug.py 2>&1 | tail -25) > gpurun_out/${TAG}_tc_vs_ffma.log 2>&1
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
(DCU_NT64=2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_nt64.json 2>gpurun_out/${TAG}_bench_nt64.err
(timeout 900 python tools/parity_report.py --frames 2048 --impls tcgen05 --out gpurun_out/${TAG}_parity_2048.json 2>&1 | tail -60) > gpurun_out/${TAG}_parity_2048.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log; tail -30 gpurun_out/${TAG}_mma_probe.log; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_nt64.json; tail -12 gpurun_out/${TAG}_parity_2048.log
