# quick GPU check: given test selection, short bench
set -x
TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_parity.py tests/test_wire.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
python bench.py --no-legs --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?
tail -5 gpurun_out/bench_$TAG.err
