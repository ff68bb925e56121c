for kb in 4096 2048 1024; do for sl in 4 8 12 16; do
SMJ_STAGE_CHUNK_KB=$kb SMJ_STAGE_SLOTS=$sl timeout 120 python tools/time_pageable.py 2>&1 | tail -1 | sed "s/^/slots=$sl /"
done; done
