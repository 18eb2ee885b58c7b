# round 2, job m: per-warp cp.async staged pair kernels — parity (staged == per-lane bits), then A/B bench
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged" 2>&1 | tail -6
one() { timeout 300 python bench.py --no-cpu "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'], d['extra']['max_neighbors'])"; }
SPHB_PAIR_MODE=1 one --warmup 60 --steps 60
SPHB_PAIR_MODE=0 one --warmup 60 --steps 60
SPHB_PAIR_MODE=1 one --warmup 20 --steps 20 --scene dam_break_10M
SPHB_PAIR_MODE=1 one --warmup 20 --steps 20 --scene fluid_drop_1M
SPHB_PAIR_MODE=0 one --warmup 20 --steps 20 --scene fluid_drop_1M
