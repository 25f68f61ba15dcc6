#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python scripts/stepper_knobs.py 1024 > gpurun_out/r2l_stepper_knobs.log 2>&1; cat gpurun_out/r2l_stepper_knobs.log | tail -30
