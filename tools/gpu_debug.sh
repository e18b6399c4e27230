#!/bin/bash
# usage (under gpurun): bash tools/gpu_debug.sh <tag> <pytest -k expression> [sanitize]
TAG=$1; KEXPR=$2; SAN=$3
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_parity_gpu.py -x -q -k "$KEXPR" > gpurun_out/${TAG}_dbg.log 2>&1
echo "plain exit $?"; tail -15 gpurun_out/${TAG}_dbg.log
if [ -n "$SAN" ]; then
  timeout 400 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_parity_gpu.py -x -q -k "$KEXPR" > gpurun_out/${TAG}_san.log 2>&1
  echo "sanitizer exit $?"; grep -v "^=========     at\|^=========         by" gpurun_out/${TAG}_san.log | head -60
fi
