#!/bin/bash
# usage (under gpurun): bash tools/gpu_bisect.sh "<pytest -k expr>" lib1.so lib2.so ...
KEXPR=$1; shift
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for L in "$@"; do
  AQSIS_B200_LIB=$PWD/aqsis_b200/_lib/$L timeout 75 python -m pytest tests/test_parity_gpu.py -x -q -k "$KEXPR" > gpurun_out/bisect_$L.log 2>&1
  echo "$L exit $?"; tail -3 gpurun_out/bisect_$L.log
done
