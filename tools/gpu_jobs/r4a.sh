# round 2 (session 3), job 4a: uploads on their own stream — whole GPU suite, bench e2e
set -x
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head -12
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; tail -c 300 gpurun_out/r4a_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r4a_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['validation']['ok'], {k: v['value'] for k, v in d['extra'].get('also', {}).items()})"
