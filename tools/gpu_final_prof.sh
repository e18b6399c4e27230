#!/bin/bash
# Evidence for profiles/: the ncu launch list of the default bench command and one full capture of the dominant kernel per
# configuration (config 4 = the headline, one band of the full-size frame; configs 2 and 3; the span filter).
TAG=$1
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --sub '' \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
bash tools/gpu_prof.sh $TAG 4:1.0:k_hide:50 2:1.0:k_hide:3 3:0.5:k_hide:3 2:1.0:k_filter_spans:3
