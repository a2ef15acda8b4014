mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_typed_keys.py -q -x 2>&1 | tail -40 > gpurun_out/r2b/typed.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b/all.log
cat gpurun_out/r2b/typed.log; tail -8 gpurun_out/r2b/all.log
