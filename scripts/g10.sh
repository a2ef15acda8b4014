mkdir -p gpurun_out/r2k
for env in "TQ_JOIN_DEBUG_POISON=1" "A=1"; do
  echo "=== $env" >> gpurun_out/r2k/torn.log
  env $env timeout 300 python scripts/diag_torn.py 8 1 >> gpurun_out/r2k/torn.log 2>&1
done
cut -c1-330 gpurun_out/r2k/torn.log | head -90
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "agg" 2>&1 | tail -5
timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | cut -c1-900
