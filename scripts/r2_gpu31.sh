#!/bin/bash
# final ncu captures of the C2 step kernels (detect, plan, force, reduce_scatter)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_detect_mask|k_force|k_reduce_scatter' -s 12 -c 4 \
   --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
   -o gpurun_out/r2z_ncu_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2z_ncu_step.log 2>&1
tail -2 gpurun_out/r2z_ncu_step.log
ncu -i gpurun_out/r2z_ncu_step.ncu-rep --page raw --csv > gpurun_out/r2z_ncu_step_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r2z_ncu_step_raw.csv "r2z: the four kernels of one C2 step (final code)" | head -60
