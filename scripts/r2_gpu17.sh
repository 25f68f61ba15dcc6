#!/bin/bash
# compute-sanitizer over the new kernels (work-list plan / force, two-cell stencil, round kernel variants)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_contact_gpu.py -m gpu -q -x -k "golden or work_list or oracle_on or edge" > gpurun_out/r2u_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2u_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_contact_gpu.py -m gpu -q -x -k "work_list or batch_matches" > gpurun_out/r2u_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2u_racecheck.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sim_gpu.py -m gpu -q -x -k "lanes or golden or traj" > gpurun_out/r2u_memcheck_sim.log 2>&1; echo "memcheck sim rc=$?"; tail -5 gpurun_out/r2u_memcheck_sim.log
