timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_cells_apart or download_between" 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
