mkdir -p gpurun_out/r2m
for env in "A=1" "TQ_AGG_NO_L2_HINT=1" "TQ_AGG_NO_FAST=1"; do
  echo "=== $env" >> gpurun_out/r2m/agg_ab.log
  env $env timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['verified']['ok'])" >> gpurun_out/r2m/agg_ab.log 2>&1
  env $env timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_agg_update -c 2 --csv python bench.py --workload agg --steps 1 --warmup 1 2>/dev/null | grep "k_agg_update" | cut -d, -f5,13- >> gpurun_out/r2m/agg_ab.log
done
cat gpurun_out/r2m/agg_ab.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2m/all.log; cat gpurun_out/r2m/all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err; cut -c1-2500 gpurun_out/r2m/bench.json; tail -5 gpurun_out/r2m/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2m/launches.csv python bench.py --kernel-only --no-secondary --steps 2 --warmup 1 > gpurun_out/r2m/ncu_bench.log 2>&1
