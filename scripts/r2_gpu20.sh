#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_contact_gpu.py tests/test_ground_gpu.py tests/test_plugin_gpu.py -m gpu -q -x 2>&1 | tail -3
for ch in 1; do
SDFC_EVAL_CHUNKS=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/r2x_chunks$ch.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2x_chunks$ch.json')); print('chunks $ch', 'step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e G q/s', round(d['e2e']['value']/1e9,2), 'value', round(d['value']/1e9,2))"
done
