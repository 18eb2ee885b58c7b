python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu > gpurun_out/b2.log 2>&1; tail -5 gpurun_out/b2.log | cut -c1-600
python -m pytest tests/test_slab_gpu.py -x -q -m gpu 2>&1 | tail -8
