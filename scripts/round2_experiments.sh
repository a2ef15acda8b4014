#!/bin/bash
# One gpurun call that (1) runs the parity tests of every round-1 experiment that has not seen a GPU yet and
# (2) A/B-times them against the defaults.  Usage on a GPU box:
#     bash scripts/round2_experiments.sh            # single GPU part
#     bash scripts/round2_experiments.sh 2          # + the 2-GPU push A/B (needs gpurun --gpus 2)
set -u
mkdir -p gpurun_out
echo "== parity of the experiments (TQ_RUN_EXPERIMENTS=1)"
TQ_RUN_EXPERIMENTS=1 timeout 900 python -m pytest tests -m gpu -q -x -k "experiment" 2>&1 | tail -15 | tee gpurun_out/r2_experiments_parity.log
echo "== join probe variants (kernel-only bench: kernel_ms = scatter + probe of 1e8 rows)"
for v in 0 1 2; do
  echo -n "TQ_JOIN_PROBE_VARIANT=$v  "
  TQ_JOIN_PROBE_VARIANT=$v timeout 300 python bench.py --kernel-only --steps 5 --warmup 3 2>/dev/null | tail -1
done | tee gpurun_out/r2_probe_variants.log
echo "== group-by: general path vs shared-memory pre-aggregation (also the partitioned variant)"
(TQ_AGG_NO_PREAGG=1 timeout 200 python scripts/agg_pre_probe.py 1000 30000 1000000
 timeout 200 python scripts/agg_pre_probe.py 1000
 TQ_AGG_PREAGG_PART=1 timeout 200 python scripts/agg_pre_probe.py 30000 1000000) 2>&1 | grep preagg | tee gpurun_out/r2_agg_preagg.log
if [ "${1:-1}" -ge 2 ]; then
  N=$1
  echo "== $N-GPU join: per-thread peer stores vs TMA bulk stores"
  for b in 0 1; do
    echo -n "TQ_PUSH_BULK=$b  "
    TQ_PUSH_BULK=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | python scripts/summ.py 2>/dev/null || true
  done | tee gpurun_out/r2_push_bulk.log
fi
