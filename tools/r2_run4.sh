timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest4.log; tail -4 gpurun_out/r2_pytest4.log
timeout 300 python tools/tputbench.py > gpurun_out/r2_tputbench.json 2> gpurun_out/r2_tputbench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_tputbench.json')); print({k:v['ms'] for k,v in d['phases'].items()})"
timeout 200 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/r2_n1_quick.json; python -c "
import json; d=json.load(open('gpurun_out/r2_n1_quick.json')); print('N=1', d['ms_per_step'], d['e2e']['ms_per_step'])"
for exb in 0 1; do
  if [ $exb = 1 ]; then export SMJ_XG_EXACT_B=1; fi
  SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n2_exb${exb}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n2_exb${exb}.json'))
print('N=2 exact_b',$exb,'ms',round(d['ms_per_step'],3),'kernels',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'), d['config']['rank0_kernels'])
PY
done
