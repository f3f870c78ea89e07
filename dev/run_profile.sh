set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "60s or per_step" 2>&1 | grep -E "parity|passed|failed" > gpurun_out/parity_numbers.log; cat gpurun_out/parity_numbers.log
python bench.py > gpurun_out/bench_dc_a.json 2> gpurun_out/bench_dc_a.err; echo rc=$?
tail -5 gpurun_out/bench_dc_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dc.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-dense-leg > gpurun_out/ncu_launch_dc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rbis_fused_kernel -s 4 -c 1 -f -o gpurun_out/prof_dc python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-dense-leg --launch-groups 1 > gpurun_out/ncu_full_dc.log 2>&1
ls -la gpurun_out/ | tail -5
