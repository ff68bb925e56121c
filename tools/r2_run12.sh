for g in 0 110 90 74; do
  SMJ_BENCH_NO_E2E=1 SMJ_XG_GRID=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n2_grid$g.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n2_grid$g.json'))
k=d['config']['rank0_kernels']
print('N=2 grid',$g,'ms',round(d['ms_per_step'],3),'kernels(serial)',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'))
PY
done
