timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "merge or mpass or join" 2>&1 | tail -3
timeout 600 python tools/time_kway.py 128000000 4 16 2>&1 | grep "k="
