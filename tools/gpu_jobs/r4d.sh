# round 2 (session 3), job 4d (2 GPUs): asynchronous slab read-back — tests, NCCL bench with the pipelined e2e
set -x
timeout 900 python -m pytest tests/test_slab_gpu.py tests/test_multi_gpu.py -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r4d_bench2.json 2> gpurun_out/r4d_bench2.err
tail -5 gpurun_out/r4d_bench2.err | cut -c1-300
python -c "
import json
d = json.loads(open('gpurun_out/r4d_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['validation']['ok'])"
