set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
bash tools/tune_run.sh --warmup 60 --steps 40 2>&1 | tee gpurun_out/r2c_tune.txt
