timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
one() { timeout 300 python bench.py --no-cpu --also "" "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'], d['extra']['max_neighbors'])"; }
one --scene dam_break_1M --warmup 5 --steps 60
one --scene dam_break_10M --warmup 5 --steps 20
one --scene fluid_drop_1M --warmup 5 --steps 20
