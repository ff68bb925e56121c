timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async or kway or tputbench" 2>&1 | tail -6 > gpurun_out/r2_pytest8.log; tail -3 gpurun_out/r2_pytest8.log
for a in 0 1; do
SMJ_SCATTER_ASYNC=$a timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 10 2>/dev/null | grep '^{' > gpurun_out/r2_n1_async$a.json; python -c "
import json; d=json.load(open('gpurun_out/r2_n1_async$a.json')); print('async $a N=1', round(d['ms_per_step'],4), {k[:24]:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items() if v['ms_per_step']>0.05})"
done
for t in 8 12 16; do SMJ_STAGE_THREADS=$t timeout 200 python tools/bench_pageable.py 2>&1 | head -1 | sed "s/^/stage threads $t: /"; done
nproc
