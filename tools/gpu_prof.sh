#!/bin/bash
# usage (under gpurun): bash tools/gpu_prof.sh <tag> <config>:<scale>:<kernel-regex>:<skip> ...
# One ncu --set full capture (with source) per item; reports land in gpurun_out/<tag>_c<config>_<n>.ncu-rep
TAG=$1; shift
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
i=0
for item in "$@"; do
  IFS=: read CFG SCALE KRE SKIP <<< "$item"
  i=$((i+1))
  ( time timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s ${SKIP:-3} -c 1 \
      -f -o gpurun_out/${TAG}_c${CFG}_${i} python bench.py --config $CFG --scale $SCALE --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub '' \
      > gpurun_out/${TAG}_c${CFG}_${i}.log 2>&1 ) 2>&1 | grep real
  echo "ncu $item exit $?"; tail -2 gpurun_out/${TAG}_c${CFG}_${i}.log | cut -c1-300
done
ls -la gpurun_out | tail -8
