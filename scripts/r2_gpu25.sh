#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_contact_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -x 2>&1 | tail -2
bash scripts/r2_ab.sh
