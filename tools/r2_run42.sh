nproc; lscpu | grep -E "NUMA|Socket" ; nvidia-smi topo -m 2>/dev/null | head -14; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; numactl -H 2>/dev/null | head -5
SMJ_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r2d_n8.err | grep "^{" > gpurun_out/r2d_n8_full.json
grep "\[bench\]" gpurun_out/r2d_n8.err | head -8
python -c "
import json; d=json.load(open('gpurun_out/r2d_n8_full.json')); print(d['ms_per_step'], d['value'], d['e2e'])"
