set -x
TAG=${1:-kb}
( for cfg in "4096 200 - 1 1" "8192 200 - 1 3" "16384 200 - 1 3" "32768 200 - 1 3"; do echo "=== $cfg"; ./dev/_build/kbench $cfg; done ) > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log
python -m pytest tests/test_gpu_group.py tests/test_gpu_synth.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
