python -m pytest tests/test_sim_gpu.py tests/test_multipart.py tests/test_ground_gpu.py tests/test_planner_gpu.py -m gpu -q -x 2>&1 | tail -2
for L in 1 2 3 4; do
SDFC_SIM_LANES=$L python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['env_steps']; print('lanes $L', round(e['ms_per_forward'],2), round(e['value']), round(e['e2e']['value']), e['rounds_per_forward'], e['outcomes'])"
done
