timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_group_test8.log; tail -3 gpurun_out/r2_group_test8.log
for n in 4 8; do
  SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n${n}_a.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n${n}_a.json'))
print('N=$n ms',round(d['ms_per_step'],3),'kernels',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'), 'sent', d['config']['exchange_bytes_sent_per_rank'], d['config']['rank0_kernels'])
PY
done
SMJ_BENCH_NO_E2E=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n8_config5.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_n8_config5.json'))
print('config5 N=8 ms',round(d['ms_per_step'],3),'value',d['value'],'kernels',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'), d['matches'], d['config']['rank0_kernels'])
PY
