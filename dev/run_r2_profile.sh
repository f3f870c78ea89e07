# round-2 profile run: GPU tests, default bench, launch list, ncu --set full of the headline kernel, of its SYN instantiation and of the configs[1] kernel
set -x
TAG=${1:-r2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?
tail -12 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rbis_fused_kernel -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_dc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs --launch-groups 1 > gpurun_out/ncu_full_${TAG}_dc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rbis_group_kernel -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_cfg1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --only-legs config1_imu_only_4096 > gpurun_out/ncu_full_${TAG}_cfg1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rbis_fused_kernel<.bool.0, .bool.1, .bool.1>' -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_syn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-legs --launch-groups 1 --e2e-steps 2 > gpurun_out/ncu_full_${TAG}_syn.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_syn.log
ls -la gpurun_out/ | tail -8
