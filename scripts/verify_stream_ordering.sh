# Regression recipe for the r1i creation-path race ("Stream ordering rule"): the test-file sequence that
# exposed it three times in a row, the SDFC_POISON=1 suite, then the full GPU suite twice. GPU box only.
for i in 1 2 3; do
  python -m pytest tests/test_multipart.py tests/test_plugin_gpu.py -m gpu -q -s 2>&1 | grep "MISMATCH\|passed\|failed\|FAILED"
done
SDFC_POISON=1 python -m pytest tests/test_contact_gpu.py tests/test_ground_gpu.py tests/test_sim_gpu.py tests/test_sdfgen.py tests/test_multipart.py tests/test_planner_gpu.py -m gpu -q 2>&1 | tail -4
for i in 1 2; do python -m pytest tests -m gpu -q -x 2>&1 | tail -3; done
