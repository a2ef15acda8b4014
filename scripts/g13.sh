mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "join" 2>&1 | tail -4
timeout 300 python bench.py --kernel-only --no-secondary --verify --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['build_ms'], d['roofline']['frac'], d['verified']['ok'])"
for ch in 4 2 8; do
  TQ_DIST_CHUNKS=$ch TQ_DIST_ONE_GPU=$([ $ch = 4 ] && echo 1 || echo 0) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ch bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n/bench2_c$ch.json 2> gpurun_out/r2n/bench2_c$ch.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2n/bench2_c$ch.json').read().strip().splitlines()[-1])
print('chunks $ch', d['ms_per_step'], d['value'], d['verified']['ok'], d.get('nccl_all_to_all_probe_exchange_ms'), d.get('one_gpu_same_job'), d['roofline']['note'])
print(d['phase_ms_rank0'])
" || tail -5 gpurun_out/r2n/bench2_c$ch.err
done
