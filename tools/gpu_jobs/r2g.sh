# round 2, job g: staged pair kernels — parity (staged == global bits), then A/B bench
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged" 2>&1 | tail -8
one() { timeout 300 python bench.py --no-cpu "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"; }
one --warmup 60 --steps 60
SPHB_PAIR_STAGE=0 one --warmup 60 --steps 60
one --warmup 20 --steps 20 --scene dam_break_10M
SPHB_PAIR_STAGE=0 one --warmup 20 --steps 20 --scene dam_break_10M
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
