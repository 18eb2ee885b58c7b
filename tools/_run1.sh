python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "refine or sixty or overflow or layout" 2>&1 | tail -3
for r in 4 5 6; do python bench.py --no-cpu --steps 20 --scene fluid_drop_1M --refine $r 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('drop refine $r', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"; done
python bench.py --no-cpu --steps 20 --refine 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dam refine 5', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
