# round 2 (session 3), job e (2 GPUs): the in-process multi-device engine over real peer copies, NCCL slab tests, both 2-GPU bench forms
set -x
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
nvidia-smi topo -m | head -6
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_slab_gpu.py -m gpu -q 2>&1 | grep -v "Warning: Particle" > gpurun_out/r3e_tests.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r3e_tests.log | head -30
timeout 900 python tools/bench_multi.py --gpus 2 --steps 30 --warmup 60 --check > gpurun_out/r3e_multi2.json 2> gpurun_out/r3e_multi2.err
tail -c 1500 gpurun_out/r3e_multi2.json; tail -3 gpurun_out/r3e_multi2.err
timeout 900 python tools/bench_multi.py --gpus 2 --scene dam_break_1M --steps 100 --warmup 60 --check > gpurun_out/r3e_multi2_1M.json 2> gpurun_out/r3e_multi2_1M.err
tail -c 1500 gpurun_out/r3e_multi2_1M.json; tail -3 gpurun_out/r3e_multi2_1M.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r3e_bench2.json 2> gpurun_out/r3e_bench2.err
tail -c 2500 gpurun_out/r3e_bench2.json; tail -5 gpurun_out/r3e_bench2.err
