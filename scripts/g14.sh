mkdir -p gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "agg" 2>&1 | tail -4
for env in "A=1" "TQ_AGG_PREAGG_PART=1" "TQ_AGG_PREAGG_PART=1 TQ_AGG_PREAGG_OLD_SCATTER=1"; do
  echo "=== $env"
  env $env timeout 300 python bench.py --workload agg --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['verified']['ok'], d['gpu_launches'])"
done
TQ_AGG_PREAGG_PART=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv python bench.py --workload agg --steps 1 --warmup 1 2>/dev/null | grep "gpu__time" | awk -F'","' '{print $5, $NF}' | cut -c1-120
timeout 300 python bench.py --kernel-only --no-secondary --verify --steps 10 --warmup 3 2>/dev/null | cut -c1-1200
