# round 2 (session 3), job b: multi-device engine + host shell + kernel-type tests, then the whole GPU suite
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_api.py tests/test_abi.py -m gpu -q 2>&1 | grep -v "Warning: Particle" > gpurun_out/r3b_tests.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r3b_tests.log | head -40
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kernel" 2>&1 > gpurun_out/r3b_kernels.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r3b_kernels.log | head -40
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning: Particle" | tail -5
