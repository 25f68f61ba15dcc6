#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python bench_sweep.py --points 1e4,1e5,1e6,1e7,1e8 --grids 64,128,256,512 --reps 5 --out gpurun_out/r2w_sweep_n1.jsonl > gpurun_out/r2w_sweep_n1.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r2w_sweep_n1.jsonl'):
    d = json.loads(ln)
    if d['points'] >= 1e6: print(d['entry'], d['points'], d['grid'], round(d['ms'], 4), 'ms', f"{d['queries_per_sec']:.3e}", 'frac', round(d['frac_of_hbm_peak'], 3), d['grid_storage'])
PY
