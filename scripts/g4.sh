mkdir -p gpurun_out/r2e
rm -f gpurun_out/r2e/diag.log
for env in "TQ_JOIN_DEBUG_SUMS=1" "TQ_JOIN_SCATTER_TILE=4096"; do
  echo "=== $env" >> gpurun_out/r2e/diag.log
  env $env timeout 120 python scripts/diag_stream.py 500000 5000000 >> gpurun_out/r2e/diag.log 2>&1
done
cat gpurun_out/r2e/diag.log
timeout 300 python bench.py --kernel-only --verify --steps 10 --warmup 3 > gpurun_out/r2e/bench_new.json 2>gpurun_out/r2e/bench_new.err; cat gpurun_out/r2e/bench_new.json; tail -3 gpurun_out/r2e/bench_new.err
TQ_JOIN_SCATTER_TILE=4096 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2e/bench_t4096.json 2>&1; cat gpurun_out/r2e/bench_t4096.json
TQ_JOIN_PART_ROWS=300000 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2e/bench_p64.json 2>&1; cat gpurun_out/r2e/bench_p64.json
TQ_JOIN_TILES_PER_CTA=16 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2e/bench_tpc16.json 2>&1; cat gpurun_out/r2e/bench_tpc16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2e/launches.csv python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2e/ncu_bench.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2e/all.log; tail -6 gpurun_out/r2e/all.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scatter_aos|k_probe_pos|k_build_part" -s 6 -c 3 -o gpurun_out/r2e/stream python bench.py --kernel-only --steps 2 --warmup 2 > gpurun_out/r2e/ncu.log 2>&1; tail -2 gpurun_out/r2e/ncu.log
