#!/bin/bash
# brick kernel: CTA width x CTAs per SM (launch bounds) x shared-memory cap per CTA, detection of 1e7 / 3e7 / 1e8 points vs 512^3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # lib cap
  SDFC_LIB_PATH=$PWD/$1 SDFC_BRICK_SMEM=$2 timeout 300 python bench_sweep.py --points 1e7,3e7,1e8 --grids 512 --reps 5 --entries detect 2>/dev/null | python -c "
import sys, json
out = []
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    out.append(f\"{d['points']:.0e}: {d['ms']:.4f} ms frac {d['frac_of_hbm_peak']:.3f} staged {d['brick_tiles']['staged']} big {d['brick_tiles']['too_large']}\")
print('$1 cap $2 |', ' | '.join(out))
"
}
run assemble-them-all_b200/libsdfcontact_b200.so 73728
for lib in $(ls tune_libs/*.so 2>/dev/null); do
  for cap in $CAPS; do run $lib $cap; done
done
