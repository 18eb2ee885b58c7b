# round 2 (session 3), job 4i (8 GPUs): strong scaling of dam_break_10M at 2 / 4 / 8 GPUs (the scene split N ways)
set -x
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --scaling strong --scene dam_break_10M --steps 60 --warmup 5 > gpurun_out/r4i_strong$n.json 2> gpurun_out/r4i_strong$n.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r4i_strong$n.json').read().strip().splitlines()[-1])
print('strong', d['n_gpus'], d['value'], d['ms_per_step'], d['config']['particles_total'], d['extra']['stage_ms_rank0'], d['extra']['ms_per_step_min'], d['extra']['ms_per_step_max'], d['e2e']['value'], d['validation']['ok'])
PY
done
