timeout 300 python tools/time_mj.py 128000000 1 2>&1 | tail -1
timeout 300 python tools/time_mj.py 32000000 10 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "merge_join or join or mpass or mpsm" 2>&1 | tail -3
