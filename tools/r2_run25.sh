timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest25.log; cat gpurun_out/r2_pytest25.log
timeout 900 python bench.py > gpurun_out/r2_bench25.json 2> gpurun_out/r2_bench25.err; tail -c 600 gpurun_out/r2_bench25.json
