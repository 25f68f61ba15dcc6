#!/bin/bash
# round-2 GPU call 11: full GPU test suite + full bench + launch list (ncu) of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2ad
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ad_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if k != 'note'}, 'launches', d['gpu_launches'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'avg_launch_ms')})
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'], 'e2e', es['e2e']['value'])
for k in ('c1_freefall', 'rotational_rollouts', 'c3_rotational_100k', 'c4_candidate_removals', 'c2_strong_scaling', 'plugin_drop_in'):
    x = d.get(k, {}); print(k, x.get('value'), x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'))
h = d['roofline_hbm_regime']; print('hbm', h['avg_launch_ms'], h['frac'], h.get('traffic_bytes_per_query'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/${T}_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/${T}_launches.csv 2>/dev/null | tail -15
