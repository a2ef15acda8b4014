mkdir -p gpurun_out/r2u
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2u/pytest.log; cat gpurun_out/r2u/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u/smoke.log 2>&1; tail -3 gpurun_out/r2u/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u/bench1.json 2> gpurun_out/r2u/bench1.err; tail -c 400 gpurun_out/r2u/bench1.err; cut -c1-260 gpurun_out/r2u/bench1.json
timeout 150 python scripts/sort_probe.py > gpurun_out/r2u/sort_probe.json 2> gpurun_out/r2u/sort_probe.err; tail -c 300 gpurun_out/r2u/sort_probe.err; cat gpurun_out/r2u/sort_probe.json
