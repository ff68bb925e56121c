timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kway or multiway or merge" 2>&1 | tail -3
SMJ_TOOL_PROFILE=1 timeout 600 python tools/time_kway.py 128000000 128 2>&1 | grep -v "^$" | tail -12
timeout 600 python tools/time_kway.py 128000000 4 8 16 32 64 256 2>&1 | grep "k="
