#!/bin/bash
# round-2 GPU call 12: launch list of the bench command including the stepper legs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2l
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 3000 --csv --log-file gpurun_out/${T}_launches_steps.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches_steps.log 2>&1
python scripts/summarize_launches.py gpurun_out/${T}_launches_steps.csv 2>/dev/null | tail -16
