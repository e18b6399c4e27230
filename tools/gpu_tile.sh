#!/bin/bash
# A/B of the static tile size on one box: bash tools/gpu_tile.sh <tag> "<config>:<scale> ..." <target1> <target2> ...  ("-" = the library's rule)
TAG=$1; ITEMS=$2; shift 2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for ROUND in 1 2; do
for item in $ITEMS; do
  IFS=: read CFG SCALE <<< "$item"
  for T in "$@"; do
    if [ "$T" = "-" ]; then unset AQH_ST_TILE; else export AQH_ST_TILE=$T; fi
    timeout 600 python bench.py --config $CFG --scale $SCALE --sub '' --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${TAG}_tile.json 2> gpurun_out/${TAG}_tile.err
    python - <<P
import json
try:
    d = json.load(open("gpurun_out/${TAG}_tile.json"))
    print("round ${ROUND} config ${CFG}@${SCALE} tile %-6s ms %.3f" % ("${T}", d["ms_per_step"]), d["stages_ms"])
except Exception as e:
    print("${T}: FAILED", e); print(open("gpurun_out/${TAG}_tile.err").read()[-600:])
P
  done
done
done
