# round 2, job p: lean staged density loop, group-unroll / occupancy variants
export SPHB_PAIR_MODE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fast or refine or overflow or sixty or ten_steps" 2>&1 | tail -4
bash tools/tune_run.sh --warmup 60 --steps 60 2>&1 | tee gpurun_out/r2p_tune.txt
