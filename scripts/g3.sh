mkdir -p gpurun_out/r2d
rm -f gpurun_out/r2d/diag.log
for env in "A=1" "TQ_JOIN_NO_TMA=1" "TQ_JOIN_SCATTER_TILE=2048" "TQ_JOIN_PART_ROWS=2000000"; do
  echo "=== $env" >> gpurun_out/r2d/diag.log
  env $env timeout 120 python scripts/diag_stream.py 500000 5000000 >> gpurun_out/r2d/diag.log 2>&1
done
cat gpurun_out/r2d/diag.log
timeout 200 ./build/ub_rank > gpurun_out/r2d/ub_rank.txt 2>&1; cat gpurun_out/r2d/ub_rank.txt
timeout 600 python -m pytest tests/test_gpu_strings.py tests/test_gpu_multikey.py tests/test_gpu_typed_keys.py -q 2>&1 | tail -15 > gpurun_out/r2d/newtests.log; tail -8 gpurun_out/r2d/newtests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scatter_aos|k_probe_pos|k_build_part" -s 6 -c 3 -o gpurun_out/r2d/stream python bench.py --kernel-only --steps 2 --warmup 2 > gpurun_out/r2d/ncu.log 2>&1; tail -3 gpurun_out/r2d/ncu.log
