#!/bin/bash
# A/B of library builds on the C2 step: bench.py --skip-env-steps for every .so under tune_libs/ and the in-tree one
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in assemble-them-all_b200/libsdfcontact_b200.so $(ls tune_libs/*.so 2>/dev/null); do
  SDFC_LIB_PATH=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-env-steps > gpurun_out/ab_$(basename $lib .so).json 2>/dev/null
  python -c "
import json,sys
d=json.load(open('gpurun_out/ab_$(basename $lib .so).json')); print('$lib', round(d['ms_per_step'],4), {k: (round(v,4) if isinstance(v, float) else '') for k,v in d['kernels_ms'].items() if k != 'note'})"
done
