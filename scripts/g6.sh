mkdir -p gpurun_out/r2g
for env in "TQ_JOIN_DEBUG_SUMS=1" "TQ_JOIN_PP_DEBUG=1" "TQ_JOIN_PP_DEBUG=2" "TQ_JOIN_PP_DEBUG=3" "TQ_JOIN_PP_VARIANT=5" "TQ_JOIN_PP_VARIANT=3"; do
  echo "=== $env" >> gpurun_out/r2g/diag.log
  env $env timeout 120 python scripts/diag_stream.py 500000 5000000 2>&1 | grep -v "missing row\|positions of\|missing sample\|missing tile" >> gpurun_out/r2g/diag.log
done
cat gpurun_out/r2g/diag.log
for t in 1024 2048 4096; do
  TQ_JOIN_SCATTER_TILE=$t TQ_JOIN_PP_VARIANT=5 timeout 300 python bench.py --kernel-only --verify --steps 8 --warmup 3 > gpurun_out/r2g/bench_t$t.json 2>gpurun_out/r2g/bench_t$t.err; echo "T$t $(cat gpurun_out/r2g/bench_t$t.json | cut -c1-420)"
done
TQ_JOIN_PP_VARIANT=5 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2g/launches.csv python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2g/ncu_bench.log 2>&1
