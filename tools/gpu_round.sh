#!/bin/bash
# One GPU-box visit: parity tests, a bench line, the ncu launch list and one full capture of k_hide.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [config]
TAG=${1:-r01}
CFG=${2:-2}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --config $CFG > gpurun_out/${TAG}_bench_c${CFG}.json 2> gpurun_out/${TAG}_bench_c${CFG}.err
echo "bench exit $?"; tail -c 3000 gpurun_out/${TAG}_bench_c${CFG}.json; tail -3 gpurun_out/${TAG}_bench_c${CFG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_c${CFG}.csv python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hide -s 3 -c 1 \
    -f -o gpurun_out/${TAG}_hide_c${CFG} python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
