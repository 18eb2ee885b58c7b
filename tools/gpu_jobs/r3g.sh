# round 2 (session 3), job g: whole GPU suite with step graphs on by default + the driver's bench line
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning: Particle" | tail -6
timeout 900 python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
tail -c 600 gpurun_out/r3g_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3g_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['extra']['stage_ms_rank0'], d['extra'].get('also', {}).get('value'))
print(d['validation']); print(d['cpu_baseline']); print(d['roofline']['frac'], d.get('roofline_issue', {}).get('frac'))
PY
