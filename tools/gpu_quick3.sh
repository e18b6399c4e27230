#!/bin/bash
# usage (under gpurun): bash tools/gpu_quick3.sh <tag> [pytest -k expr | "none"] [configs...]
# GPU tests, then resident-only bench lines (no e2e, no CPU baseline, no sub-records) of the given configs.
TAG=$1; KEXPR=$2; shift 2
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
if [ "$KEXPR" != "none" ]; then
  if [ -n "$KEXPR" ] && [ "$KEXPR" != "all" ]; then
    ( time timeout 1500 python -m pytest tests -q -m gpu --durations=5 -k "$KEXPR" ) > gpurun_out/${TAG}_pytest.log 2>&1
  else
    ( time timeout 1500 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/${TAG}_pytest.log 2>&1
  fi
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${TAG}_pytest.log | head -30
fi
for CFG in "$@"; do
  timeout 600 python bench.py --config $CFG --sub '' --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/${TAG}_q_c${CFG}.json 2> gpurun_out/${TAG}_q_c${CFG}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${TAG}_q_c${CFG}.json"))
    print("config ${CFG}: ms_per_step %.3f" % d["ms_per_step"], d["stages_ms"], "launches", d["gpu_launches"], "bands", d.get("bands"))
except Exception as e:
    print("config ${CFG}: FAILED", e); print(open("gpurun_out/${TAG}_q_c${CFG}.err").read()[-1500:])
P
done
