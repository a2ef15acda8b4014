mkdir -p gpurun_out/r2l
nvidia-smi -L > gpurun_out/r2l/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r2l/dist_check.log 2>&1; tail -25 gpurun_out/r2l/dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l/bench2.json 2> gpurun_out/r2l/bench2.err; cut -c1-1500 gpurun_out/r2l/bench2.json; tail -15 gpurun_out/r2l/bench2.err
