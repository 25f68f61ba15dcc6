#!/bin/bash
# round-2 GPU call 1: full GPU test suite, live peaks, brick-vs-records A/B in the HBM regime, ncu of both
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.log
python - > gpurun_out/r2a_peaks.log 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from assemble_them_all_b200 import contact
ctx = contact.Context(0)
for _ in range(3):
    print(ctx.measure_peaks())
PY
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
for lay in records bricks; do for ord in none morton kd; do
  timeout 600 python bench_sweep.py --points 3e7 --grids 512 --entries detect --layout $lay --order $ord --reps 7 >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
done; done
timeout 600 python bench_sweep.py --points 1e8 --grids 512 --entries detect --layout bricks --order kd --reps 5 >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
timeout 600 python bench_sweep.py --points 1e8 --grids 512 --entries detect --layout records --order morton --reps 5 >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
for lay in records bricks; do
  ord=kd; [ $lay = records ] && ord=morton
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_detect -s 4 -c 1 -o gpurun_out/r2a_ncu_$lay \
     python bench_sweep.py --points 3e7 --grids 512 --entries detect --layout $lay --order $ord --reps 2 > gpurun_out/r2a_ncu_$lay.log 2>&1
done
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_peaks.log; cat gpurun_out/r2a_ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['grid_storage'], d['point_order'], d['points'], round(d['ms'],4), 'ms', round(d['frac_of_hbm_peak'],3))
"
