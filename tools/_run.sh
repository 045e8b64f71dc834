# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is `bash tools/collect_profiles.sh <tag> [parity frames]`
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
for i in 1 2 3; do python bench.py --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['e2e_python']['value']))"; done
