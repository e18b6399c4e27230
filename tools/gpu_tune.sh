#!/bin/bash
# usage (under gpurun): bash tools/gpu_tune.sh <tag> <config> <scale> <tune1> <tune2> ...   (each tune = an AQH_TUNE value, "-" = defaults)
TAG=$1; CFG=$2; SCALE=$3; shift 3
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for T in "$@"; do
  if [ "$T" = "-" ]; then unset AQH_TUNE; else export AQH_TUNE=$T; fi
  timeout 600 python bench.py --config $CFG --scale $SCALE --sub '' --no-e2e --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/${TAG}_tune.json 2> gpurun_out/${TAG}_tune.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${TAG}_tune.json"))
    print("config ${CFG} scale ${SCALE} tune ${T}: ms %.3f" % d["ms_per_step"], d["stages_ms"])
except Exception as e:
    print("tune ${T}: FAILED", e); print(open("gpurun_out/${TAG}_tune.err").read()[-800:])
P
done
