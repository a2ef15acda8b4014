mkdir -p gpurun_out/r2o
for ch in 4 2 3; do
  TQ_DIST_CHUNKS=$ch TQ_DIST_ONE_GPU=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$ch bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o/bench2_c$ch.json 2> gpurun_out/r2o/bench2_c$ch.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2o/bench2_c$ch.json').read().strip().splitlines()[-1])
print('chunks $ch', d['ms_per_step'], d['value'], d['verified']['ok'], d['roofline']['note'][-80:])
print(d['phase_ms_rank0'])
" || tail -5 gpurun_out/r2o/bench2_c$ch.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_check.py 2>&1 | grep "OK\|Error\|error" | head
