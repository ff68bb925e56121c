set -x
python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_py_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/prof_join.py > gpurun_out/plain_join.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_join python tools/prof_join.py > gpurun_out/ncu_join.log 2>&1
python tools/prof_kway.py > gpurun_out/plain_kway.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_kw -o gpurun_out/r02_kway python tools/prof_kway.py > gpurun_out/ncu_kway.log 2>&1
cat gpurun_out/plain_kway.log | tail -12
ls -la gpurun_out/*.ncu-rep
