#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> <config> [extra bench args]
TAG=$1; N=$2; CFG=$3; shift 3
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt; nvidia-smi topo -m >> gpurun_out/${TAG}_smi.txt 2>&1
python tools/pcie_probe.py > gpurun_out/${TAG}_pcie.txt 2>&1; cat gpurun_out/${TAG}_pcie.txt
for n in 1 $N; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  ( time timeout 400 $L bench.py --gpus $n --config $CFG --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_c${CFG}_n${n}.json 2> gpurun_out/${TAG}_bench_c${CFG}_n${n}.err ) 2>&1 | tail -3
  echo "n=$n exit $?"; tail -c 1800 gpurun_out/${TAG}_bench_c${CFG}_n${n}.json; tail -5 gpurun_out/${TAG}_bench_c${CFG}_n${n}.err
done
