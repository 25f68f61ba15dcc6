#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/c2step.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from assemble_them_all_b200 import contact, scenes
ctx = contact.Context(0)
rs = scenes.c2_planner_rollouts(n_envs=1024)
sim, batch, *keep = rs.upload(ctx)
sim.reset(rs.q0, rs.qdot0)
for b in range(len(rs.bodies)):
    sim.set_body_external_force(b, rs.f_ext[:, b])
sim.forward(60)
ctx.synchronize()
print(sim.stats(), batch.skin_stats())
PY
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_detect_mask' --launch-skip 200 --launch-count 1 -o gpurun_out/r2y_ncu_detect_step python /tmp/c2step.py > gpurun_out/r2y_ncu_detect_step.log 2>&1
tail -2 gpurun_out/r2y_ncu_detect_step.log
ncu -i gpurun_out/r2y_ncu_detect_step.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__waves_per_multiprocessor'):
        print(k, d.get(k))
"
