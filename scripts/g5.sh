mkdir -p gpurun_out/r2f
python scripts/diag_default_inner.py > gpurun_out/r2f/definner.log 2>&1; cat gpurun_out/r2f/definner.log
TQ_JOIN_DEBUG_SUMS=1 timeout 120 python scripts/diag_stream.py 500000 5000000 > gpurun_out/r2f/diag.log 2>&1; tail -6 gpurun_out/r2f/diag.log
for v in 0 1 2 3 4 5; do
  TQ_JOIN_PP_VARIANT=$v timeout 300 python bench.py --kernel-only --verify --steps 8 --warmup 3 > gpurun_out/r2f/bench_v$v.json 2>gpurun_out/r2f/bench_v$v.err; echo "v$v $(cat gpurun_out/r2f/bench_v$v.json | cut -c1-330)"
done
TQ_JOIN_PP_VARIANT=3 TQ_JOIN_TILES_PER_CTA=4 timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 > gpurun_out/r2f/bench_v3_tpc4.json 2>&1; cat gpurun_out/r2f/bench_v3_tpc4.json | cut -c1-200
TQ_JOIN_PP_VARIANT=3 TQ_JOIN_TILES_PER_CTA=32 timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 > gpurun_out/r2f/bench_v3_tpc32.json 2>&1; cat gpurun_out/r2f/bench_v3_tpc32.json | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2f/all.log; tail -8 gpurun_out/r2f/all.log
