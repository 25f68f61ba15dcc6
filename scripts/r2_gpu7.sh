#!/bin/bash
# round-2 GPU call 7: work-list force kernel (k_force_plan + persistent k_force): tests, bench, G=2, ncu; C1 tail launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_tests.log
tail -4 gpurun_out/r2g_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['kernels_ms'])
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'], 'e2e', es['e2e']['value'])
for k in ('c1_freefall', 'rotational_rollouts', 'c3_rotational_100k', 'c4_candidate_removals', 'c2_strong_scaling', 'plugin_drop_in'):
    x = d.get(k, {}); print(k, x.get('value'), x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'))
PY
SDFC_FORCE_G=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2g_bench_g2.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2g_bench_g2.json')); print('G=2', d['ms_per_step'], d['kernels_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 6 -c 2 \
   --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
   -o gpurun_out/r2g_ncu_force python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2g_ncu_force.log 2>&1
ncu -i gpurun_out/r2g_ncu_force.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum','dram__bytes_read.sum','dram__bytes_write.sum','sass__inst_executed_local_loads','sass__inst_executed_local_stores','smsp__inst_executed.sum'):
        print(k, d.get(k))
"
# C1: plain run, then the launch list of a window of straggler rounds
timeout 300 python scripts/c1_tail.py 1024 100 > gpurun_out/r2g_c1_tail.log 2>&1; cat gpurun_out/r2g_c1_tail.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 --launch-count 400 --csv --log-file gpurun_out/r2g_c1_tail_launches.csv python scripts/c1_tail.py 1024 100 > gpurun_out/r2g_c1_tail_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2g_c1_tail_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size') if 'Grid Size' in hdr else None
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in agg.items(): print(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us max={max(v)/1e3:8.2f} us")
PY
