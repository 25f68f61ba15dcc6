#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2ab_bench_short.json 2> gpurun_out/r2ab_bench_short.err; echo rc=$?; tail -2 gpurun_out/r2ab_bench_short.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ab_bench_short.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'])
print(json.dumps(d['roofline'])[:600])
PY
