timeout 300 python tools/time_mj.py 32000000 10 2>&1 | tail -2
SMJ_TOOL_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_mj_tile -o gpurun_out/r02_mjtile python tools/time_mj.py 128000000 1 > gpurun_out/ncu_mj.log 2>&1
tail -3 gpurun_out/ncu_mj.log
