for n in 8 4; do
  SMJ_BENCH_NO_E2E=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/r2c_n$n.err | grep '^{' > gpurun_out/r2c_n$n.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2c_n$n.json'))
k=d['config']['rank0_kernels']
print('N=$n ms',round(d['ms_per_step'],3),'kernels(serial)',round(d['config']['rank0_kernel_ms_per_step'],3), {kk[:16]:round(v,3) for kk,v in k.items()}, d['matches'], d['checksum'])
PY
done
