for c in 8192 4096 3072 2048 1536; do SMJ_STAGE_CHUNK_KB=$c timeout 120 python tools/time_pageable.py 2>&1 | tail -1; done
for c in 2048 4096; do SMJ_STAGE_CHUNK_KB=$c timeout 120 python tools/time_pageable.py 2>&1 | tail -1; done
