# scratch command file for `gpurun --gpus 8 -- 'bash tools/_run.sh'`
TAG=r2n
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 2>&1 | grep -E "DIST_CHECK|Error|error" | head -5)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench8.err | tail -1) > gpurun_out/${TAG}_bench_8gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_8gpu.json")); print("8gpu", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "py", round(d["e2e_python"]["value"]), "one_batch", d.get("one_batch"), d["clocks"])
PY
tail -3 gpurun_out/${TAG}_bench8.err
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 2>/dev/null | tail -1) > gpurun_out/${TAG}_bench_4gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_4gpu.json")); print("4gpu", round(d["value"]), "fps", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), "py", round(d["e2e_python"]["value"]), "one_batch", d.get("one_batch"))
PY
