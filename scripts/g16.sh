mkdir -p gpurun_out/r2q
for env in "A=1" "TQ_AGG_PREAGG_PART=0"; do
  env $env timeout 600 python scripts/agg_pre_probe.py 3000 30000 300000 1000000 5000000 2>&1 | sed "s/^/$env  /" >> gpurun_out/r2q/agg_groups.log
done
cat gpurun_out/r2q/agg_groups.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2q/all.log; cat gpurun_out/r2q/all.log
timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('agg', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['verified']['ok'], d['gpu_launches'])"
timeout 300 python bench.py --kernel-only --no-secondary --verify --steps 10 --warmup 3 2>/dev/null | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scatter_aos|k_probe_pos|k_build_part" -s 6 -c 4 -o gpurun_out/r2q/join python bench.py --kernel-only --no-secondary --steps 2 --warmup 1 > gpurun_out/r2q/ncu_join.log 2>&1; tail -3 gpurun_out/r2q/ncu_join.log
