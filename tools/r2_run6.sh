timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q -k "key8b or kway or multiway or small_sort or group or gpus or device_sort" 2>&1 | tail -15 > gpurun_out/r2_pytest6.log; tail -4 gpurun_out/r2_pytest6.log
timeout 300 python tools/tputbench.py > gpurun_out/r2_tputbench2.json 2> gpurun_out/r2_tputbench2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_tputbench2.json')); print({k:v['ms'] for k,v in d['phases'].items()})"
for tile in 8192 12288 16384; do
  SMJ_BENCH_NO_E2E=1 SMJ_XG_TILE=$tile timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n2_tile${tile}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n2_tile${tile}.json'))
k=d['config']['rank0_kernels']
print('N=2 tile',$tile,'ms',round(d['ms_per_step'],3),'kernels',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'), 'passA', [v for kk,v in k.items() if 'onesweep_unstable' in kk], 'launches/step', d['gpu_launches']/2/10)
PY
done
