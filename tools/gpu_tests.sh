#!/bin/bash
# every GPU test (or the ones matching $2): bash tools/gpu_tests.sh <tag> [pytest -k expression]
TAG=${1:-t}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  ( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 -k "$2" ) > gpurun_out/${TAG}_pytest.log 2>&1
else
  ( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 ) > gpurun_out/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${TAG}_pytest.log | head -40
