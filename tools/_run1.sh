python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or full_size" 2>&1 | tail -2
python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dam', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
