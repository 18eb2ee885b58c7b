python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or full_size" 2>&1 | tail -3
bash tools/tune_run.sh
