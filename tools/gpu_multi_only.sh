#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi_only.sh <tag> <N> <config> [extra bench args]   (the N-rank leg only)
TAG=$1; N=$2; CFG=$3; shift 3
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_c${CFG}_n${N}.json 2> gpurun_out/${TAG}_bench_c${CFG}_n${N}.err ) 2>&1 | grep real
echo "n=$N exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_c${CFG}_n${N}.json; tail -4 gpurun_out/${TAG}_bench_c${CFG}_n${N}.err | cut -c1-300
