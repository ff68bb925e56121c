timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "merge_join or join or mpass or mpsm or bulk" 2>&1 | tail -5
timeout 300 python tools/time_mj.py 128000000 1 2>&1 | tail -2
timeout 300 python tools/time_mj.py 32000000 10 2>&1 | tail -2
