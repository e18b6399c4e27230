#!/bin/bash
# usage (under gpurun): bash tools/gpu_mbtest.sh <tag> -- MB/DoF parity tests first (bounded), then the whole suite, then bench config 3
TAG=$1
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_parity_gpu.py -x -q -k "mb_dof or motion or dof" > gpurun_out/${TAG}_mb.log 2>&1
rc=$?; echo "mb tests exit $rc"; tail -6 gpurun_out/${TAG}_mb.log
[ $rc -ne 0 ] && exit 1
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --config 3 --no-cpu-baseline --steps 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_c3.json; tail -3 gpurun_out/${TAG}_bench_c3.err
