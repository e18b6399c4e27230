#!/bin/bash
# usage (under gpurun): bash tools/gpu_bench_cfg.sh <tag> <config> [extra bench args]
TAG=$1; CFG=$2; shift 2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 1500 python bench.py --config $CFG "$@" > gpurun_out/${TAG}_bench_c${CFG}.json 2> gpurun_out/${TAG}_bench_c${CFG}.err ) 2>&1 | tail -4
echo "bench c$CFG exit $?"; tail -c 2500 gpurun_out/${TAG}_bench_c${CFG}.json; tail -5 gpurun_out/${TAG}_bench_c${CFG}.err
