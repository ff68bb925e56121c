for n in 4 8; do for g in 1000 110 92; do
  SMJ_BENCH_NO_E2E=1 SMJ_XG_GRID=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n${n}_grid$g.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_n${n}_grid$g.json'))
print('N=$n grid',$g,'ms',round(d['ms_per_step'],3),'kernels(serial)',round(d['config']['rank0_kernel_ms_per_step'],3), d['config'].get('group_path'))
PY
done; done
SMJ_BENCH_NO_E2E=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n8_config5_b.json
python -c "
import json; d=json.load(open('gpurun_out/r2_n8_config5_b.json')); print('config5', round(d['ms_per_step'],3), d['value'], d['matches'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/r2_n8_full.json
python -c "
import json; d=json.load(open('gpurun_out/r2_n8_full.json')); print('N=8 full', round(d['ms_per_step'],3), 'e2e', d['e2e'])"
