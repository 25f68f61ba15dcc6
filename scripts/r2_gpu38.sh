#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/c2s.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from assemble_them_all_b200 import contact, scenes
ctx = contact.Context(0)
rs = scenes.c2_planner_rollouts(n_envs=128)
sim, batch, *keep = rs.upload(ctx)
sim.reset(rs.q0, rs.qdot0)
for b in range(len(rs.bodies)):
    sim.set_body_external_force(b, rs.f_ext[:, b])
sim.forward(100)
ctx.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 600 --csv --log-file gpurun_out/r2ad_launches_128.csv python /tmp/c2s.py > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2ad_launches_128.csv 2>/dev/null | tail -8
