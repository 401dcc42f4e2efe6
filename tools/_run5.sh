set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_scale2.json 2> gpurun_out/r02_scale2.err
tail -5 gpurun_out/r02_scale2.err
python -m pytest tests/test_gpu_abi5.py -q -k two_gpus 2>&1 | tail -3
