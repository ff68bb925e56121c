timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest41.log; cat gpurun_out/r2_pytest41.log
timeout 900 python bench.py > gpurun_out/r02e_bench_default_line.json 2> gpurun_out/r02e_bench.err; tail -c 200 gpurun_out/r02e_bench_default_line.json
python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02e_bench_py_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
