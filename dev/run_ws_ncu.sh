set -x
WS_MAPPINGS=32 ncu --set full --clock-control none --import-source on -k regex:rbis_ws_kernel -s 3 -c 1 -f -o gpurun_out/prof_ws_imu python dev/ws_bench.py 4096:imu > gpurun_out/ncu_ws_imu.log 2>&1
tail -3 gpurun_out/ncu_ws_imu.log
