# round 2 (session 3), job x (8 GPUs): dam_break_100M with the fixed validation (int64 ownership checksum; the single-GPU reference run keeps grid refine 4), N = 4 weak
set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --scene dam_break_100M --scaling strong --steps 20 --warmup 5 > gpurun_out/r3x_bench8_100M.json 2> gpurun_out/r3x_bench8_100M.err
tail -c 300 gpurun_out/r3x_bench8_100M.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3x_bench8_100M.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['particles_total'], d['extra']['stage_ms_rank0'], d['validation'])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r3x_bench4.json 2> gpurun_out/r3x_bench4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3x_bench4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['particles_total'], d['extra']['stage_ms_rank0'], d['validation']['ok'])
PY
