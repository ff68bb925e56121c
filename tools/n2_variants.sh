timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_group_test2.log; tail -4 gpurun_out/r2_group_test2.log
for tile in 8192 16384; do for ovl in 0 1; do
  SMJ_BENCH_NO_E2E=1 SMJ_XG_TILE=$tile SMJ_XG_OVERLAP=$ovl timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n2_t${tile}_o${ovl}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n2_t${tile}_o${ovl}.json'))
k=d['config']['rank0_kernels']
print('tile',$tile,'overlap',$ovl,'ms',round(d['ms_per_step'],3),'kernels',round(d['config']['rank0_kernel_ms_per_step'],3),'passA',[v for kk,v in k.items() if 'onesweep_unstable' in kk], d['config'].get('group_path'))
PY
done; done
