#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_r2_gpu.py tests/test_contact_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench_sweep.py --points 1e6,1e7,3e7,1e8 --grids 256,512 --reps 5 --entries detect --out gpurun_out/r2aa_sweep_detect.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print(d['entry'], d['points'], d['grid'], round(d['ms'], 4), 'ms', f\"{d['queries_per_sec']:.3e}\", 'frac', round(d['frac_of_hbm_peak'], 3), d['grid_storage'], d['brick_tiles'])
"
