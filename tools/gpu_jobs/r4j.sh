# round 2 (session 3), job 4j (8 GPUs): the one-process engine (sphb_create_multi) on the strong-scaled dam_break_10M, 4 and 8 GPUs
set -x
for n in 4 8; do
  timeout 600 python tools/bench_multi.py --gpus $n --scaling strong --scene dam_break_10M --steps 100 --warmup 60 > gpurun_out/r4j_multi_strong$n.json 2> gpurun_out/r4j_multi_strong$n.err
  python -c "
import json
d = json.loads(open('gpurun_out/r4j_multi_strong$n.json').read().strip().splitlines()[-1])
print('multi strong', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3), d['config']['owned'], d['config']['halo_copies'])"
done
