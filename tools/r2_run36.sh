timeout 300 python tools/time_lsd.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_algos.py 128000000 16 2>&1 | grep -v "^    "
