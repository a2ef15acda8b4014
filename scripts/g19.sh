mkdir -p gpurun_out/r2t
timeout 400 python -m pytest tests/test_gpu_chunk_codec.py tests/test_gpu_final_agg.py tests/test_gpu_sort_merge.py -m gpu -q --durations=8 2>&1 | tail -60 > gpurun_out/r2t/new.log; cat gpurun_out/r2t/new.log
timeout 500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_chunk_codec.py --deselect tests/test_gpu_final_agg.py --deselect tests/test_gpu_sort_merge.py 2>&1 | tail -12 > gpurun_out/r2t/rest.log; cat gpurun_out/r2t/rest.log
