set -x
TAG=${1:-kb}
( for cfg in "65536 200 - 1 3" "113664 200 - 1 3"; do echo "=== $cfg"; ./dev/_build/kbench $cfg; ./dev/_build/kbench $cfg; done ) > gpurun_out/kbench_$TAG.log 2>&1
cat gpurun_out/kbench_$TAG.log
