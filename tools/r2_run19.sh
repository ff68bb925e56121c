timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kway or multiway or merge" 2>&1 | tail -5
SMJ_TOOL_PROFILE=1 timeout 600 python tools/time_kway.py 128000000 128 4 16 2>&1 | tail -60
