#!/bin/bash
# usage: scripts/dist_bench.sh N [extra bench.py args]   — runs bench.py on N GPUs of this box under torchrun
N=${1:-2}; shift
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "${PORT:-29533}" bench.py --gpus "$N" "$@"
