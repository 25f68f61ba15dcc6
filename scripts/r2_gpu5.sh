#!/bin/bash
# round-2 GPU call 5: warp-per-pair planner force kernel (tests, bench, occupancy variants, flop census), plug-in batching
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
tail -4 gpurun_out/r2e_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2e_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['kernels_ms'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'avg_launch_ms')}, 'detect', {k: d['roofline_detect_c2'][k] for k in ('achieved', 'peak', 'frac')})
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'], es['rounds_per_forward'], es['gpu_launches_per_forward'], 'e2e', es['e2e']['value'])
for k in ('c1_freefall', 'rotational_rollouts', 'c3_rotational_100k', 'c4_candidate_removals', 'c2_strong_scaling', 'plugin_drop_in'):
    x = d.get(k, {}); print(k, x.get('value'), x.get('ms_per_forward'), x.get('rounds_per_forward'), x.get('converged'), x.get('stock_cpu'), x.get('speedup_vs_stock_1core'))
h = d['roofline_hbm_regime']; print('hbm', h['avg_launch_ms'], h['frac'], h['cell_records_morton']['avg_launch_ms'], h['cell_records_morton']['frac'], h['brick_tiles'])
PY
for lib in tune_libs/libnf2_mb2.so tune_libs/libnf2_mb4.so; do
  SDFC_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2e_$(basename $lib .so).json 2>/dev/null
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2e_$(basename $lib .so).json')); print('$lib', d['ms_per_step'], d['kernels_ms'])"
done
# executed FP64 instruction census + traffic of the planner force kernel and of detection on C2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_nf -s 6 -c 1 \
   --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
   -o gpurun_out/r2e_ncu_force python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2e_ncu_force.log 2>&1
ncu -i gpurun_out/r2e_ncu_force.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum','dram__bytes_read.sum','dram__bytes_write.sum','sass__inst_executed_local_loads','sass__inst_executed_local_stores'):
        print(k, d.get(k))
"
