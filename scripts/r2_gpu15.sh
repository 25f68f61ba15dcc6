#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tests/drivers/redmax_rollout.py oracle/_ref/b200 /tmp/b.npz --nverts 10000 --dx 0.05 --steps 100
python tests/drivers/redmax_rollout.py oracle/_ref/b200 /tmp/b.npz --nverts 10000 --dx 0.05 --steps 100
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 600 --csv --log-file gpurun_out/r2r_plugin_launches.csv python tests/drivers/redmax_rollout.py oracle/_ref/b200 /tmp/b.npz --nverts 10000 --dx 0.05 --steps 100 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2r_plugin_launches.csv 2>/dev/null | tail -12
