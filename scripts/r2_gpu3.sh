#!/bin/bash
# round-2 GPU call 3: warp-per-environment stepper (tests + bench), brick kernel after LDS / all-inside changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['kernels_ms'])
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'], 'e2e', es['e2e']['value'])
for k in ('c1_freefall', 'rotational_rollouts', 'c4_candidate_removals'):
    x = d[k]; print(k, x['value'], x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'))
print(d['roofline_hbm_regime'])
PY
rm -f gpurun_out/r2c_ab.jsonl
for cap in 55296 73728; do
  SDFC_BRICK_SMEM=$cap timeout 600 python bench_sweep.py --points 3e7,1e8 --grids 512 --entries detect --layout bricks --order kd --reps 7 >> gpurun_out/r2c_ab.jsonl 2>> gpurun_out/r2c_ab.err
done
cat gpurun_out/r2c_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['grid_storage'], d['point_order'], d['points'], d['brick_smem'], round(d['ms'],4), 'ms', round(d['frac_of_hbm_peak'],3), d['brick_tiles'])
"
