# round 2 (session 3), job y: whole GPU suite (final state of the session) and the driver's default bench command
set -x
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r3y_bench.json 2> gpurun_out/r3y_bench.err ) 2>&1 | grep real
python -c "
import json
d = json.loads(open('gpurun_out/r3y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['steps'], d['extra']['stage_ms_rank0'], d['validation']['ok'], d['e2e']['value'], {k: v['value'] for k, v in d['extra'].get('also', {}).items()}, d['clocks'])"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -5 | cut -c1-600
