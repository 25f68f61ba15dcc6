#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_planner_gpu.py -m gpu -q -x --durations=3 2>&1 | tail -8
