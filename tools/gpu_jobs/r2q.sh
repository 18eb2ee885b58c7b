# round 2, job q: shared DensityWalker in both kernels: staged == per-lane bits again, tolerance tests, variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
bash tools/tune_run.sh --warmup 60 --steps 60 2>&1 | tee gpurun_out/r2q_tune.txt
SPHB_PAIR_MODE=1 python bench.py --no-cpu --warmup 60 --steps 60 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('staged', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
python bench.py --no-cpu --warmup 20 --steps 20 --scene dam_break_10M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10M lane', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
SPHB_PAIR_MODE=1 python bench.py --no-cpu --warmup 20 --steps 20 --scene dam_break_10M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10M staged', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
