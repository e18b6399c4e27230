#!/bin/bash
# The two lines the driver takes at round end, on one box: bash tools/gpu_bench_final.sh <tag>
TAG=$1
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
( time timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err ) 2>&1 | grep real
( time timeout 1500 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real
python - <<P
import json
for f in ("gpurun_out/${TAG}_bench_reference_arm.json", "gpurun_out/${TAG}_bench_default.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms", d.get("ms_per_step"), "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "cold", (d.get("e2e_cold") or {}).get("value"))
        for k, v in (d.get("sub") or {}).items():
            print("   ", k, v.get("ms_per_step"), v.get("value"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
P
tail -3 gpurun_out/${TAG}_bench.err
