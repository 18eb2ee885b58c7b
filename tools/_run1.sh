set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or full_size or sparse or ten_steps" 2>&1 | tail -15
for v in 0 2; do python bench.py --no-cpu --steps 10 --pair-kernel $v 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"; done
