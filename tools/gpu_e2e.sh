#!/bin/bash
# usage (under gpurun): bash tools/gpu_e2e.sh <tag> <config>...   resident + e2e (warm and cold) of the given configs, no CPU baseline
TAG=$1; shift
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for CFG in "$@"; do
  timeout 900 python bench.py --config $CFG --sub '' --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/${TAG}_e2e_c${CFG}.json 2> gpurun_out/${TAG}_e2e_c${CFG}.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${TAG}_e2e_c${CFG}.json"))
    print("config ${CFG}: resident %.2f ms" % d["ms_per_step"], d["stages_ms"], "| e2e %.2f ms" % d["e2e"]["ms_per_step"], d["e2e"]["last_step_ms"], "| cold %.2f ms" % d["e2e_cold"]["ms_per_step"], "prepare", d["e2e_cold"]["prepare_ms"], "bands", d.get("bands"))
except Exception as e:
    print("config ${CFG}: FAILED", e); print(open("gpurun_out/${TAG}_e2e_c${CFG}.err").read()[-1500:])
P
done
