timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -3
SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2c_n2.err | grep '^{' > gpurun_out/r2c_n2.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2c_n2.json'))
k=d['config']['rank0_kernels']
print('N=2 ms',round(d['ms_per_step'],3),'kernels(serial)',round(d['config']['rank0_kernel_ms_per_step'],3), {kk[:16]:round(v,3) for kk,v in k.items()}, d['matches'], d['checksum'])
PY
