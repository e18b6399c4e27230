#!/bin/bash
# usage (under gpurun): bash tools/gpu_quick.sh <tag> [configs...]  -- parity tests + one bench line per config
TAG=${1:-q}; shift
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
( time timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1 ) 2>&1 | grep real
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for CFG in "$@"; do
  EXTRA=""; [ "$CFG" != "2" ] && EXTRA="--no-cpu-baseline"
  ( time timeout 420 python bench.py --config $CFG $EXTRA > gpurun_out/${TAG}_bench_c${CFG}.json 2> gpurun_out/${TAG}_bench_c${CFG}.err ) 2>&1 | grep real
  echo "bench c$CFG exit $?"; tail -c 2600 gpurun_out/${TAG}_bench_c${CFG}.json; tail -3 gpurun_out/${TAG}_bench_c${CFG}.err
done
