#!/bin/bash
# round-2 GPU call 2: brick kernel A/B over the shared-memory budget (0 = global gathers only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2b_ab.jsonl
timeout 300 python -m pytest tests/test_parity_r2_gpu.py -q -x > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.log
for cap in 0 73728 112640 204800; do
  SDFC_BRICK_SMEM=$cap timeout 600 python bench_sweep.py --points 3e7,1e8 --grids 512 --entries detect --layout bricks --order kd --reps 7 >> gpurun_out/r2b_ab.jsonl 2>> gpurun_out/r2b_ab.err
done
SDFC_BRICK_SMEM=73728 timeout 600 python bench_sweep.py --points 3e7 --grids 512 --entries detect --layout bricks --order morton --reps 7 >> gpurun_out/r2b_ab.jsonl 2>> gpurun_out/r2b_ab.err
SDFC_BRICK_SMEM=112640 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_detect -s 4 -c 1 -o gpurun_out/r2b_ncu_bricks110 \
     python bench_sweep.py --points 3e7 --grids 512 --entries detect --layout bricks --order kd --reps 2 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_tests.log
cat gpurun_out/r2b_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['grid_storage'], d['point_order'], d['points'], d['brick_smem'], round(d['ms'],4), 'ms', round(d['frac_of_hbm_peak'],3), d['brick_tiles'])
"
