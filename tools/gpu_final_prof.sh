#!/bin/bash
# launch list + full captures of the two dominant kernels on config 2 (for profiles/)
TAG=$1
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
bash tools/gpu_prof.sh $TAG 2:1.0:k_hide:3 2:1.0:k_filter_spans:1
