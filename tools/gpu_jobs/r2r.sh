# round 2, job r: lean staged force pass: parity (all modes identical bits), variants (mode 1 = both passes staged)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
export SPHB_PAIR_MODE=1
bash tools/tune_run.sh --warmup 60 --steps 60 2>&1 | tee gpurun_out/r2r_tune.txt
python bench.py --no-cpu --warmup 20 --steps 20 --scene dam_break_10M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10M staged', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
SPHB_PAIR_MODE=2 python bench.py --no-cpu --warmup 20 --steps 20 --scene dam_break_10M 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10M mixed', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
SPHB_PAIR_MODE=2 python bench.py --no-cpu --warmup 60 --steps 60 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1M mixed', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
