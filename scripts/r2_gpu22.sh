#!/bin/bash
# A/B of library builds on the stepper (forward(100) x 1024 rollouts, both joint types)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in assemble-them-all_b200/libsdfcontact_b200.so $(ls tune_libs/*.so 2>/dev/null); do
  echo "--- $lib"
  SDFC_LIB_PATH=$PWD/$lib timeout 300 python scripts/stepper_knobs.py 1024 gv 2>&1 | grep "GV=1"
done
