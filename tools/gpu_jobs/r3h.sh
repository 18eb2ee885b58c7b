# round 2 (session 3), job h: first step at which two slabs differ from the single context at 2 x 10.6 M (fast mode), one GPU
set -x
timeout 1200 python tools/debug/multi_bisect.py dam_break_10M 0,0 90 2>&1 | tail -20
timeout 900 python bench.py --steps 30 --warmup 5 --also "" > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; tail -c 300 gpurun_out/r3h_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r3h_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['validation'])"
