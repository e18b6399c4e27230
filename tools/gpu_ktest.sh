#!/bin/bash
# usage (under gpurun): bash tools/gpu_ktest.sh <tag> "<pytest -k expr>"  -- bounded subset first, then the whole gpu suite
TAG=$1; KEXPR=$2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -k "$KEXPR" > gpurun_out/${TAG}_k.log 2>&1
rc=$?; echo "subset exit $rc"; tail -25 gpurun_out/${TAG}_k.log | cut -c1-400
[ $rc -ne 0 ] && exit 1
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
