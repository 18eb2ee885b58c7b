# round 2 (session 3), job 4f (8 GPUs): the driver's scaling command at N = 8 on the final code (pipelined e2e in the slab path)
set -x
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r4f_bench8.json 2> gpurun_out/r4f_bench8.err ) 2>&1 | grep real
tail -c 400 gpurun_out/r4f_bench8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r4f_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['particles_total'], d['e2e'], d['extra']['stage_ms_rank0'], d['extra']['ms_per_step_min'], d['extra']['ms_per_step_max'], d['validation']['ok'])
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 ) 2>&1 | tail -6 | cut -c1-300
