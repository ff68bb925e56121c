timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "mpsm or join" 2>&1 | tail -3
timeout 600 python tools/time_algos.py 128000000 16 2>&1 | tail -4
SMJ_MPSM_PER_RUN=1 timeout 600 python tools/time_algos.py 128000000 16 2>&1 | grep "^mpsm"
