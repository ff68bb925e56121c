set -x
timeout 600 python tools/tputbench.py > gpurun_out/r02b_tputbench.json 2> gpurun_out/r02b_tputbench.err; tail -c 400 gpurun_out/r02b_tputbench.json
timeout 900 python bench.py > gpurun_out/r02b_bench_default_line.json 2> gpurun_out/r02b_bench.err
python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_bench_py_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r02b_bench_reference_arm.json | head -c 600
