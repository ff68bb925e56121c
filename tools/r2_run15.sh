# bulk-store output of the scatter kernels (SMJ_SCATTER_BULK=1) and the persistent L2-prefetch variant
mkdir -p gpurun_out
for v in "" "SMJ_SCATTER_BULK=1" "SMJ_SCATTER_PERSIST=1"; do
  echo "== $v"; env $v timeout 300 python tools/time_join.py 128000000 10 2>&1 | tail -2
done
SMJ_SCATTER_BULK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "join or sort or partition or msd or scatter" 2>&1 | tail -5
