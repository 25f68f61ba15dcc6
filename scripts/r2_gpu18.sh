#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plugin_gpu.py tests/test_contact_gpu.py tests/test_ground_gpu.py tests/test_parity_r2_gpu.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2 3; do python tests/drivers/redmax_rollout.py oracle/_ref/b200 /tmp/b.npz --nverts 10000 --dx 0.05 --steps 100 --reps 3; done
SDFC_RUN_GRAPH=0 python tests/drivers/redmax_rollout.py oracle/_ref/b200 /tmp/b.npz --nverts 10000 --dx 0.05 --steps 100 --reps 3
bash scripts/r2_ab.sh
