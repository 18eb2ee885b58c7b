# round 2, job u (2 GPUs): bench line at N = 2 (weak, 2 x 10M) incl. the slab validation; NCCL parity test
set -x
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2u_bench2.json 2> gpurun_out/r2u_bench2.err
tail -c 2500 gpurun_out/r2u_bench2.json; tail -5 gpurun_out/r2u_bench2.err
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q 2>&1 | tail -3
