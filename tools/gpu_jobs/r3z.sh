# round 2 (session 3), job z: pipelined read-back (sphb_download_begin / _end) — test, e2e of the bench line
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "download_begin" 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err; tail -c 300 gpurun_out/r3z_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r3z_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])"
