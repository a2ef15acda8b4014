mkdir -p gpurun_out/r2j
for env in "A=1" "TQ_JOIN_DEBUG_SUMS=1" "TQ_JOIN_PP_DEBUG=1" "TQ_JOIN_PP_DEBUG=2" "TQ_JOIN_NO_TMA=1" "TQ_JOIN_SCATTER_TILE=1024" "TQ_JOIN_PP_VARIANT=1"; do
  echo "=== $env" >> gpurun_out/r2j/torn.log
  env $env timeout 200 python scripts/diag_torn.py 4 1 >> gpurun_out/r2j/torn.log 2>&1
done
echo "=== no marker" >> gpurun_out/r2j/torn.log
timeout 200 python scripts/diag_torn.py 6 0 >> gpurun_out/r2j/torn.log 2>&1
grep -v "tq debug" gpurun_out/r2j/torn.log | head -150
timeout 600 python -m pytest tests/test_gpu_multikey.py -q -x -k default_inner 2>&1 | tail -5
