#!/bin/bash
# the GPU suite with fresh device allocations poisoned (0xFF): no result may depend on memory the library never wrote
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SDFC_POISON=1 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_tests_poison.log 2>&1; echo "poison rc=$?"; tail -3 gpurun_out/r2z_tests_poison.log
SDFC_RUN_GRAPH=0 SDFC_SIM_GRAPH=0 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_tests_nograph.log 2>&1; echo "nograph rc=$?"; tail -3 gpurun_out/r2z_tests_nograph.log
