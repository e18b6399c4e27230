#!/bin/bash
# A/B of experimental library builds on ONE box: bash tools/gpu_ab.sh <tag> "<config>:<scale> ..." <lib1> <lib2> ...
# ("-" = the in-tree library); two rounds, interleaved, resident frames only.
TAG=$1; ITEMS=$2; shift 2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for ROUND in 1 2; do
for item in $ITEMS; do
  IFS=: read CFG SCALE <<< "$item"
  for LIBF in "$@"; do
    if [ "$LIBF" = "-" ]; then unset AQSIS_B200_LIB; else export AQSIS_B200_LIB=$PWD/$LIBF; fi
    timeout 600 python bench.py --config $CFG --scale $SCALE --sub '' --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${TAG}_ab.json 2> gpurun_out/${TAG}_ab.err
    python - <<P
import json
try:
    d = json.load(open("gpurun_out/${TAG}_ab.json"))
    print("round ${ROUND} config ${CFG}@${SCALE} %-28s ms %.3f" % ("${LIBF}", d["ms_per_step"]), d["stages_ms"])
except Exception as e:
    print("${LIBF}: FAILED", e); print(open("gpurun_out/${TAG}_ab.err").read()[-600:])
P
  done
done
done
