mkdir -p gpurun_out/r2i
timeout 300 python scripts/diag_fail.py default > gpurun_out/r2i/diag_default.log 2>&1; tail -40 gpurun_out/r2i/diag_default.log
timeout 300 python scripts/diag_fail.py shapes > gpurun_out/r2i/diag_shapes.log 2>&1; tail -30 gpurun_out/r2i/diag_shapes.log
TQ_JOIN_OLD_FAST=1 timeout 300 python scripts/diag_fail.py shapes > gpurun_out/r2i/diag_shapes_old.log 2>&1; tail -12 gpurun_out/r2i/diag_shapes_old.log
timeout 600 python -m pytest tests/test_gpu_expr_program.py -q -x 2>&1 | tail -25 > gpurun_out/r2i/expr.log; cat gpurun_out/r2i/expr.log
timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | cut -c1-400
TQ_AGG_NO_FAST=1 timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_agg_update -c 2 -o gpurun_out/r2i/agg python bench.py --workload agg --steps 1 --warmup 1 > gpurun_out/r2i/ncu_agg.log 2>&1; tail -3 gpurun_out/r2i/ncu_agg.log
