timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "sort or merge or kway" 2>&1 | tail -3
timeout 600 python tools/time_kway.py 128000000 4 8 16 32 128 2>&1 | grep "k="
timeout 300 python tools/time_sort.py 1000000000 2>&1 | grep "^n \|bucket_sort"
timeout 300 python tools/time_sort.py 128000000 2>&1 | grep "^n \|bucket_sort"
