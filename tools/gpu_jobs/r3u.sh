set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -k "lanes" 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -12
bash tools/gpu_jobs/r3t.sh 2>&1 | grep scene
