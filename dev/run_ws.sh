set -x
TAG=${1:-ws}
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
RBIS_DEBUG_ATTR=1 timeout 600 python dev/ws_bench.py > gpurun_out/ws_bench_$TAG.log 2>&1; cat gpurun_out/ws_bench_$TAG.log
