# scratch: non-default switches still produce oracle-identical results (64 frames each)
TAG=r2r
mkdir -p gpurun_out
for cfg in "A=1" "DCU_TC_PAIR=0" "DCU_FLAT=0" "DCU_FUSE_UP=0" "DCU_NT64=0" "DCU_NT64=1" "DCU_SEG=1" "DCU_SEG=2" "DCU_FUSE_FIRST=1" "DCU_WRES=0" "DCU_WRES_UP=0" "DCU_SLICE_MINOR=0" "DCU_DEVICE_COUNT=0" "DCU_ARG_HEADS=0" "DCU_GRAPH=0" "DCU_PDL=0" "DCU_CONV_IMPL=ffma"; do
  echo "== $cfg: $( (env $cfg timeout 300 python tools/parity_report.py --frames 64 --impls tcgen05 --out gpurun_out/${TAG}_p.json 2>&1 | grep -E '^tcgen05_f16x2|Error|error' | cut -c1-330) )"
done
