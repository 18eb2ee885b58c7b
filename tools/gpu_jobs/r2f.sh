timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu --warmup 20 --steps 20 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra'], d['gpu_launches'])"
python bench.py --no-cpu --warmup 20 --steps 20 --scene dam_break_10M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra'], d['gpu_launches'])"
