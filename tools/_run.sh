# scratch command file for `gpurun -- 'bash tools/_run.sh'`; the full evidence run is `bash tools/collect_profiles.sh <tag> [parity frames]`
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python bench.py 2>/dev/null | tail -1
