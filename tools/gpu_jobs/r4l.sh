# round 2 (session 3), job 4l (8 GPUs): the one-process engine with per-device worker threads — strong-scaled and weak-scaled dam break; tests over real peers
set -x
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -5
for n in 4 8; do
  timeout 600 python tools/bench_multi.py --gpus $n --scaling strong --scene dam_break_10M --steps 100 --warmup 60 > gpurun_out/r4l_multi_strong$n.json 2> gpurun_out/r4l_multi_strong$n.err
  python -c "
import json
d = json.loads(open('gpurun_out/r4l_multi_strong$n.json').read().strip().splitlines()[-1])
print('multi strong', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3))"
done
timeout 600 python tools/bench_multi.py --gpus 8 --steps 30 --warmup 60 > gpurun_out/r4l_multi_weak8.json 2> gpurun_out/r4l_multi_weak8.err
python -c "
import json
d = json.loads(open('gpurun_out/r4l_multi_weak8.json').read().strip().splitlines()[-1])
print('multi weak', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 3))"
