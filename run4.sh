cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --config 2 --no-cpu-baseline > gpurun_out/bench_c2b.json 2> gpurun_out/bench_c2b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2b.json'))
print({k:d[k] for k in ('value','ms_per_step','stages_ms','e2e','counters')})
PY
tail -3 gpurun_out/bench_c2b.err
