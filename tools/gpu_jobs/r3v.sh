# round 2 (session 3), job v: whole GPU suite after the lanes-per-particle fix (explicit default of 1 lane), bench line
set -x
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head -12
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err; tail -c 300 gpurun_out/r3v_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r3v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'], d['validation']['ok'], d['e2e']['value'], {k: v['value'] for k, v in d['extra'].get('also', {}).items()})"
