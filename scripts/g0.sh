mkdir -p gpurun_out/r2a
TQ_RUN_EXPERIMENTS=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a/tests_exp.log
for v in 0 1 2; do TQ_JOIN_PROBE_VARIANT=$v timeout 200 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2a/bench_var$v.json 2>gpurun_out/r2a/bench_var$v.err; done
for pr in 300000 600000; do TQ_JOIN_PART_ROWS=$pr timeout 200 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2a/bench_part$pr.json 2>&1; done
TQ_JOIN_TILES_PER_CTA=4 timeout 200 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2a/bench_tpc4.json 2>&1
TQ_JOIN_TILES_PER_CTA=16 timeout 200 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2a/bench_tpc16.json 2>&1
nvidia-smi > gpurun_out/r2a/smi.txt
cat gpurun_out/r2a/tests_exp.log | tail -15; cat gpurun_out/r2a/bench_*.json
