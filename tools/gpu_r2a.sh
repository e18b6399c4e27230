#!/bin/bash
# round 2, visit A: every GPU test, then the default bench line (config 4 + sub-records) and the reference arm
TAG=${1:-r2a}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; free -g >> gpurun_out/${TAG}_smi.txt
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
echo "ref exit $?"; tail -c 600 gpurun_out/${TAG}_ref.json; tail -3 gpurun_out/${TAG}_ref.err
