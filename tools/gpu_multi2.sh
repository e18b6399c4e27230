#!/bin/bash
# multi-GPU bench lines: bash tools/gpu_multi2.sh <tag> <ngpus> <config> [extra bench args]
TAG=$1; N=$2; CFG=$3; shift 3
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --config $CFG "$@" ) > gpurun_out/${TAG}_bench_c${CFG}_n${N}.json 2> gpurun_out/${TAG}_bench_c${CFG}_n${N}.err
echo "bench c$CFG n$N exit $?"; tail -c 2500 gpurun_out/${TAG}_bench_c${CFG}_n${N}.json; tail -5 gpurun_out/${TAG}_bench_c${CFG}_n${N}.err
