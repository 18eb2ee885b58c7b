# round 2 (session 3), job 4g (2 GPUs): the driver's exact scaling command (default steps / warmup) after the validation fix
set -x
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r4g_bench2.json 2> gpurun_out/r4g_bench2.err ) 2>&1 | grep real
python -c "
import json
d = json.loads(open('gpurun_out/r4g_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['e2e']['value'], d['validation'])"
