timeout 300 python tools/time_lsd.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or partition" 2>&1 | tail -3
