mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "stream or fast_kernel or 5e6 or large or partition" 2>&1 | tail -30 > gpurun_out/r2c/stream.log
tail -5 gpurun_out/r2c/stream.log
timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c/bench_new.json 2>gpurun_out/r2c/bench_new.err; cat gpurun_out/r2c/bench_new.json; tail -3 gpurun_out/r2c/bench_new.err
TQ_JOIN_OLD_FAST=1 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c/bench_old.json 2>&1; cat gpurun_out/r2c/bench_old.json
TQ_JOIN_SCATTER_TILE=2048 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c/bench_t2048.json 2>&1; cat gpurun_out/r2c/bench_t2048.json
TQ_JOIN_PART_ROWS=300000 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c/bench_p64.json 2>&1; cat gpurun_out/r2c/bench_p64.json
TQ_JOIN_PART_ROWS=300000 TQ_JOIN_SCATTER_TILE=2048 timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c/bench_p64_t2048.json 2>&1; cat gpurun_out/r2c/bench_p64_t2048.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2c/ncu_bench.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2c/all.log; tail -6 gpurun_out/r2c/all.log
