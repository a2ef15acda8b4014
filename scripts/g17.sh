mkdir -p gpurun_out/r2r
nvidia-smi -L | wc -l
TQ_DIST_LEAN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2r/bench8.json 2> gpurun_out/r2r/bench8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2r/bench8.json').read().strip().splitlines()[-1])
print('N=8', d['ms_per_step'], d['value'], d['verified'], d['roofline']['note'][-90:])
print(d['phase_ms_rank0'])
" || tail -20 gpurun_out/r2r/bench8.err
