#!/bin/bash
# round-2 GPU call 14: A/B of tune libs, parity tests with both layouts, detect kernel ncu, 1-GPU sweep at 512^3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash scripts/r2_ab.sh
timeout 900 python -m pytest tests/test_parity_r2_gpu.py tests/test_contact_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench_sweep.py --points 1e7,1e8 --grids 256,512 --reps 5 --out gpurun_out/r2q_sweep_n1_big.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print(d['entry'], d['points'], d['grid'], round(d['ms'], 4), 'ms', f\"{d['queries_per_sec']:.3e}\", 'frac', round(d['frac_of_hbm_peak'], 3), d['grid_storage'])
"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_detect_mask$' -s 3 -c 1 -o gpurun_out/r2q_ncu_detect python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2q_ncu_detect.log 2>&1
ncu -i gpurun_out/r2q_ncu_detect.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum'):
        print(k, d.get(k))
"
