# round 2, job i: staged kernels with suspend hint, deeper rings, split force records: parity + variants
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged" 2>&1 | tail -4
bash tools/tune_run.sh --warmup 60 --steps 60 2>&1 | tee gpurun_out/r2i_tune.txt
SPHB_PAIR_STAGE=0 python bench.py --no-cpu --warmup 60 --steps 60 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('global', d['value'], d['ms_per_step'], d['extra']['stage_ms_rank0'])"
