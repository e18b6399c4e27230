#!/bin/bash
# usage (under gpurun, with a --phase-timing build in the tree): bash tools/gpu_phase.sh <tag> <config>:<scale> ...
TAG=$1; shift
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for item in "$@"; do
  IFS=: read CFG SCALE <<< "$item"
  timeout 600 python bench.py --config $CFG --scale $SCALE --sub '' --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2> gpurun_out/${TAG}_phase_c${CFG}.txt
  echo "== config $CFG scale $SCALE"; grep "aqh phase" gpurun_out/${TAG}_phase_c${CFG}.txt | tail -7
done
