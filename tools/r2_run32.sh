timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest32.log; cat gpurun_out/r2_pytest32.log
timeout 600 python tools/tputbench.py > gpurun_out/r02c_tputbench.json 2> gpurun_out/r02c_tputbench.err; tail -c 300 gpurun_out/r02c_tputbench.json
timeout 600 python tools/time_kway.py 128000000 4 8 16 32 64 128 256 2>&1 | grep "k=" > gpurun_out/r02c_kway_by_k.txt; cat gpurun_out/r02c_kway_by_k.txt
