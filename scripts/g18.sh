mkdir -p gpurun_out/r2s
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2s/pytest.log; cat gpurun_out/r2s/pytest.log
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/r2s/bench1.json 2> gpurun_out/r2s/bench1.err; tail -c 600 gpurun_out/r2s/bench1.err; cut -c1-300 gpurun_out/r2s/bench1.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s/launches.csv python bench.py --kernel-only --no-secondary --steps 2 --warmup 1 > gpurun_out/r2s/ncu.log 2>&1; tail -2 gpurun_out/r2s/ncu.log
