# N=2: does the NVLink-bound exchange pass gain from bulk (TMA) stores?
for v in "SMJ_X=0" "SMJ_SCATTER_BULK=1"; do
  env $v SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2_n2_bulk.err | grep '^{' > gpurun_out/r2_n2_bulk_$v.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n2_bulk_$v.json'))
k=d['config']['rank0_kernels']
print('$v','ms',round(d['ms_per_step'],3),'kernels',round(d['config']['rank0_kernel_ms_per_step'],3),'passA',[v for kk,v in k.items() if 'onesweep_unstable' in kk], d['config'].get('group_path'), d['matches'])
PY
done
