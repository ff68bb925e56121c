timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort" 2>&1 | tail -3
timeout 300 python tools/time_sort.py 1000000000 2>&1 | tail -14
timeout 300 python tools/time_sort.py 128000000 2>&1 | tail -14
