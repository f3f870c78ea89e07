set -x
ncu --set full --clock-control none --import-source on -k regex:rbis_group_kernel -s 1 -c 1 -f -o gpurun_out/prof_g8lean ./dev/_build/kbench 8192 200 g8_w16 1 3 > gpurun_out/ncu_g8lean.log 2>&1
tail -3 gpurun_out/ncu_g8lean.log
