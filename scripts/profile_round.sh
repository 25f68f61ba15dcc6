set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r1i_tests.log
python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err
python bench.py --impl reference > gpurun_out/r1i_bench_reference.json 2>> gpurun_out/r1i_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_detect_mask|k_force|k_sim_post|k_sim_pre" --launch-skip 40 -c 8 -o gpurun_out/r1i_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
