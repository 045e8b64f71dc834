python tools/bench_configs.py --iters 5 --out gpurun_out/r1c_configs.json > gpurun_out/r1c_configs.log 2>&1; tail -50 gpurun_out/r1c_configs.log
python tools/bench_single_frame.py --cpu > gpurun_out/r1c_single_frame.json 2> gpurun_out/r1c_single_frame.err; cat gpurun_out/r1c_single_frame.json
