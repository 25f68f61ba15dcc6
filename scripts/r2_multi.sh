#!/bin/bash
# round-2 multi-GPU call: bench.py and the C5 sweep with points sharded over N ranks (N = number of visible GPUs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r2p
if [ -z "$SKIP_BENCH" ]; then
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench_n$N.json'))
print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
es = d['env_steps']; print('env_steps', es['value'], es['ms_per_forward'])
for k in ('c2_strong_scaling', 'c4_candidate_removals'):
    x = d.get(k, {}); print(k, x.get('value'), x.get('ms_per_stage'), x.get('ms_per_forward'), x.get('scaling'))
PY
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench_sweep.py --points ${POINTS:-1e4,1e5,1e6,1e7,1e8} --grids ${GRIDS:-64,128,256,512} --reps 5 --out gpurun_out/${T}_sweep_n$N.jsonl > gpurun_out/${T}_sweep_n$N.log 2>&1; echo "sweep rc=$?"
python - <<PY
import json
for ln in open('gpurun_out/${T}_sweep_n$N.jsonl'):
    d = json.loads(ln)
    print(d['entry'], d['points'], d['grid'], d['n_gpus'], round(d['ms'], 4), 'ms', f"{d['queries_per_sec']:.3e}", 'q/s frac', round(d['frac_of_hbm_peak'], 3), d['grid_storage'])
PY
