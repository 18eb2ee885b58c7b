python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast" 2>&1 | tail -2
bash tools/tune_run.sh
