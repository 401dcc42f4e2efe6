python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_scale8.json 2> gpurun_out/r02_scale8.err
tail -3 gpurun_out/r02_scale8.err
