#!/bin/bash
# Run ON THE GPU BOX: compute-sanitizer over the small end-to-end invocation (__graft_entry__.smoke(): 4 frames through every kernel of
# the pipeline incl. the CTA-pair tcgen05 kernels with their mbarrier / TMA / tensor-memory pipelines).
#   memcheck  : out-of-bounds / misaligned global + shared accesses
#   racecheck : shared-memory hazards between the producer / MMA / epilogue warps (the async-proxy writes of TMA and tcgen05 are
#               ordered by mbarriers, which racecheck understands for cp.async.bulk)
#   synccheck : invalid barrier / cluster-barrier usage
# Usage: bash tools/sanitize.sh <tag>    -> gpurun_out/<tag>_sanitize_{memcheck,racecheck,synccheck}.log
TAG=${1:-rX}
mkdir -p gpurun_out
export DCU_GRAPH=0          # kernel-by-kernel launches (the sanitizer does not see inside graph replays as well)
for tool in memcheck racecheck synccheck; do
  (timeout 420 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25) > gpurun_out/${TAG}_sanitize_$tool.log 2>&1
  echo "== $tool: exit $? =="; tail -4 gpurun_out/${TAG}_sanitize_$tool.log
done
