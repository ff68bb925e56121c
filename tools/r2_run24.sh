nproc; lscpu | grep -E "Model name|Socket|NUMA|L3|Thread" 
for wc in 0 1; do for t in 4 8 12 16; do
SMJ_STAGE_WC=$wc SMJ_STAGE_THREADS=$t timeout 120 python tools/time_pageable.py 2>&1 | tail -1 | sed "s/^/wc=$wc /"
done; done
