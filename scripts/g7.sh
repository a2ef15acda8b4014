mkdir -p gpurun_out/r2h
for i in 1 2 3 4 5 6; do
  for env in "A=1" "TQ_JOIN_PP_VARIANT=2" "TQ_JOIN_SCATTER_TILE=1024"; do
    env $env timeout 120 python scripts/diag_stream.py 500000 5000000 2>&1 | grep "dup ids" | sed "s/^/$env  /" >> gpurun_out/r2h/stress.log
  done
done
env TQ_JOIN_DEBUG_SUMS=1 timeout 120 python scripts/diag_stream.py 3000000 20000000 2>&1 | grep -v "missing row\|positions of\|missing sample\|missing tile" >> gpurun_out/r2h/stress.log
cat gpurun_out/r2h/stress.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2h/all.log; tail -10 gpurun_out/r2h/all.log
timeout 300 python bench.py --kernel-only --verify --steps 10 --warmup 3 > gpurun_out/r2h/bench.json 2>gpurun_out/r2h/bench.err; cat gpurun_out/r2h/bench.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2h/launches.csv python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2h/ncu_bench.log 2>&1
timeout 300 python bench.py --workload agg --steps 5 --warmup 3 > gpurun_out/r2h/bench_agg.json 2>gpurun_out/r2h/bench_agg.err; cat gpurun_out/r2h/bench_agg.json | cut -c1-600; tail -3 gpurun_out/r2h/bench_agg.err
