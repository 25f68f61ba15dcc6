#!/bin/bash
# round-2 GPU call 8: rolled column loop in the planner force kernel (instruction-cache footprint)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 900 python -m pytest tests/test_contact_gpu.py tests/test_parity_r2_gpu.py tests/test_sim_gpu.py tests/test_plugin_gpu.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
for lib in assemble-them-all_b200/libsdfcontact_b200.so; do
  SDFC_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/${T}_$(basename $lib .so).json 2>/dev/null
  python -c "
import json,sys
d=json.load(open('gpurun_out/${T}_$(basename $lib .so).json')); print('$lib', d['ms_per_step'], d['kernels_ms'], d['roofline']['frac'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_force$' -s 3 -c 1 \
   --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
   -o gpurun_out/${T}_ncu_force python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/${T}_ncu_force.log 2>&1
ncu -i gpurun_out/${T}_ncu_force.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum','dram__bytes_read.sum','dram__bytes_write.sum','sass__inst_executed_local_loads','sass__inst_executed_local_stores','smsp__inst_executed.sum'):
        print(k, d.get(k))
"
