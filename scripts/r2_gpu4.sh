#!/bin/bash
# round-2 GPU call 4: fused + graphed stepper rounds (tests, bench)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2d_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['kernels_ms'])
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'], 'e2e', es['e2e']['value'])
for k in ('c1_freefall', 'rotational_rollouts', 'c4_candidate_removals'):
    x = d[k]; print(k, x['value'], x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'))
PY
SDFC_SIM_GRAPH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_nograph.json 2> gpurun_out/r2d_bench_nograph.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2d_bench_nograph.json'))
es = d['env_steps']; print('NOGRAPH env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'])
for k in ('c1_freefall', 'rotational_rollouts', 'c4_candidate_removals'):
    x = d[k]; print('NOGRAPH', k, x['value'], x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'))
PY
