python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dam', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
python bench.py --no-cpu --steps 5 --scene fluid_drop_1M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('drop', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
