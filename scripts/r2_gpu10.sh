#!/bin/bash
# round-2 GPU call 10: source-level profile of the tail round kernel on C1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_sim_round' --launch-skip 8000 --launch-count 1 -o gpurun_out/r2k_ncu_round python scripts/c1_tail.py 1024 100 > gpurun_out/r2k_ncu_round.log 2>&1
tail -3 gpurun_out/r2k_ncu_round.log
ncu -i gpurun_out/r2k_ncu_round.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys, csv
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ('Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','sass__inst_executed_local_loads','sass__inst_executed_local_stores'):
        print(k, d.get(k))
"
