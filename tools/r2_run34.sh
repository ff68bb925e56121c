SMJ_TOOL_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_mj_tile -o gpurun_out/r02_mjtile_tma18 python tools/time_mj.py 128000000 1 > gpurun_out/ncu_mj3.log 2>&1
tail -2 gpurun_out/ncu_mj3.log
timeout 600 python tools/time_algos.py 128000000 16 2>&1 | grep -v "^    "
