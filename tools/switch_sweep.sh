#!/bin/bash
# Run ON THE GPU BOX: every A/B switch of DESIGN.md 8b against the oracle (tools/parity_report.py, 64 frames each) -> one table.
# Usage: bash tools/switch_sweep.sh [out file]
OUT=${1:-gpurun_out/switch_sweep.txt}
mkdir -p "$(dirname "$OUT")"
echo "Every A/B switch of DESIGN.md 8b against the oracle (tools/parity_report.py --frames 64 --impls tcgen05, seed 1), one B200." > "$OUT"
echo "Columns: corners, differences in {kept set, ids, raw pixel, refined position} (all must be 0), error maxima." >> "$OUT"
printf "%-22s %8s %5s %4s %7s %5s %11s %11s %11s\n" switch corners kept ids raw_px heat "max|dloc|" "max|dids|" "max|dheat|" >> "$OUT"
for sw in "" DCU_TC_PAIR=0 DCU_FLAT=0 DCU_FUSE_UP=0 DCU_NT64=0 DCU_NT64=1 DCU_NT64=2 DCU_SEG=1 DCU_SEG=2 DCU_FUSE_FIRST=1 DCU_WRES=0 \
          DCU_WRES_UP=0 DCU_SLICE_MINOR=0 DCU_DEVICE_COUNT=0 DCU_ARG_HEADS=0 DCU_GRAPH=0 DCU_PDL=0 DCU_SMALL_SLICES=0 DCU_OVERLAP_FIRST=0; do
  env $sw timeout 300 python tools/parity_report.py --frames 64 --impls tcgen05 --out /tmp/sweep.json > /tmp/sweep.log 2>&1
  python - "$sw" >> "$OUT" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/sweep.json"))
    r = d["impls"]["tcgen05_f16x2"]
    print("%-22s %8d %5d %4d %7d %5d %11.4f %11.4f %11.2e" % (sys.argv[1] or "(defaults)", r["K"], r["kept_set"], r["ids"], r["raw_px"],
                                                          r["heat_flip"], r["max_abs_dloc"], r["max_abs_dids"], r["max_abs_dheat"]))
except Exception as e:
    print("%-22s FAILED: %s" % (sys.argv[1] or "(defaults)", e))
PY
done
cat "$OUT"
