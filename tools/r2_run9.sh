timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest9.log; tail -4 gpurun_out/r2_pytest9.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; tail -3 gpurun_out/r2_bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench9.json'))
print('N=1', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],2), 'pinned', round(d['e2e']['pinned_host_memory']['ms_per_step'],2), 'cpu', d['cpu_baseline'])
for k,v in d['configs'].items():
    print(k, v.get('ms_per_step'), v.get('error'), {a:round(b['ms'],2) for a,b in v.get('thetas',{}).items()})
PY
for t in 12 16; do SMJ_STAGE_THREADS=$t timeout 200 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stage threads $t: pageable e2e', round(d['e2e']['pageable_host_memory']['ms_per_step'],2))"; done
