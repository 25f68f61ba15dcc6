#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_contact_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python scripts/stepper_knobs.py 1024 gv 2>&1 | tee gpurun_out/r2t_gv.log | tail -10
