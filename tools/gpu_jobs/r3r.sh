# round 2 (session 3), job r (8 GPUs): the driver's scaling command at N = 8 (weak, 8 x 10.6 M), wall time included; dam_break_100M; in-process multi engine on 8 GPUs
set -x
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
free -g | head -2
date +%s
time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r3r_bench8.json 2> gpurun_out/r3r_bench8.err
tail -c 400 gpurun_out/r3r_bench8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3r_bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['particles_total'], d['e2e']['value'], d['extra']['stage_ms_rank0'], d['validation']['ok'])
PY
time timeout 900 python tools/bench_multi.py --gpus 8 --steps 30 --warmup 60 > gpurun_out/r3r_multi8.json 2> gpurun_out/r3r_multi8.err
tail -c 900 gpurun_out/r3r_multi8.json; tail -3 gpurun_out/r3r_multi8.err
time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --scene dam_break_100M --scaling strong --steps 20 --warmup 5 > gpurun_out/r3r_bench8_100M.json 2> gpurun_out/r3r_bench8_100M.err
tail -c 300 gpurun_out/r3r_bench8_100M.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3r_bench8_100M.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['particles_total'], d['extra']['stage_ms_rank0'], d['validation']['ok'])
PY
