#!/bin/bash
# round-2 GPU call 9: tail variant of the round kernel (CTA per environment): parity tests, C1 tail timing for team sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2k
timeout 1500 python -m pytest tests/test_sim_gpu.py tests/test_configs_gpu.py tests/test_planner_gpu.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
echo "--- tail off"; SDFC_SIM_TAIL=0 timeout 300 python scripts/c1_tail.py 1024 100 2>&1 | tail -1
for lib in assemble-them-all_b200/libsdfcontact_b200.so tune_libs/libtail64.so tune_libs/libtail256.so; do
  echo "--- $lib"; SDFC_LIB_PATH=$PWD/$lib timeout 300 python scripts/c1_tail.py 1024 100 2>&1 | tail -1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 --launch-count 300 --csv --log-file gpurun_out/${T}_c1_tail_launches.csv python scripts/c1_tail.py 1024 100 > gpurun_out/${T}_c1_tail_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2k_c1_tail_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in agg.items(): print(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us max={max(v)/1e3:8.2f} us")
PY
