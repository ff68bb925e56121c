# multi-GPU record after the bulk-store exchange became the default (run with gpurun --gpus 8)
for n in 2 4 8; do
  SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r2b_n$n.err | grep '^{' > gpurun_out/r2b_n$n.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_n$n.json'))
k=d['config']['rank0_kernels']
print('N=$n ms',round(d['ms_per_step'],3),'kernels(serial)',round(d['config']['rank0_kernel_ms_per_step'],3),'passA',[round(v,3) for kk,v in k.items() if 'onesweep_unstable' in kk], d['config'].get('group_path'), d['matches'], d['checksum'])
PY
done
SMJ_XG_BULK=0 SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2b_n8_nobulk.json
python -c "
import json; d=json.load(open('gpurun_out/r2b_n8_nobulk.json')); print('N=8 without bulk stores', round(d['ms_per_step'],3))"
SMJ_BENCH_NO_E2E=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2b_n8_config5.json
python -c "
import json; d=json.load(open('gpurun_out/r2b_n8_config5.json')); print('config5', round(d['ms_per_step'],3), d['value'], d['matches'])"
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -3
