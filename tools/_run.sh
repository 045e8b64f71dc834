# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is `bash tools/collect_profiles.sh <tag> [parity frames]`
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
for r in 1 2; do
for v in old new; do
  cp _ab/lib_$v.so deepcharuco_b200/libdeepcharuco_b200.so
  echo "== $v"; python bench.py --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
done; done
