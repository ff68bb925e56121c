SMJ_TOOL_PROFILE=1 timeout 600 python tools/time_kway.py 128000000 128 2>&1 | tail -12
cat > /tmp/prof_kw.py <<'PY'
import os, sys, ctypes as C
sys.argv = ["x", "128000000", "128"]
PY
SMJ_TOOL_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kw_tile_count -c 1 -o gpurun_out/r02_kwcount python tools/time_kway.py 128000000 128 > gpurun_out/ncu_kwcount.log 2>&1
tail -3 gpurun_out/ncu_kwcount.log
