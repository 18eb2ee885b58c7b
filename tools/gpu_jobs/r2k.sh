# round 2, job k: two-particles-per-thread pair kernels — parity, then A/B bench against mode 0 and register caps
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
bash tools/tune_run.sh --warmup 60 --steps 60 2>&1 | tee gpurun_out/r2k_tune.txt
one() { timeout 300 python bench.py --no-cpu "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'], d['extra']['max_neighbors'])"; }
SPHB_PAIR_MODE=0 one --warmup 60 --steps 60
one --warmup 20 --steps 20 --scene dam_break_10M
one --warmup 20 --steps 20 --scene fluid_drop_1M
SPHB_PAIR_MODE=0 one --warmup 20 --steps 20 --scene fluid_drop_1M
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
