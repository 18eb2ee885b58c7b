# round 2 (session 3), job a: the multi-device engine behind the C ABI (sphb_create_multi) on one GPU (slabs on ordinal 0) + host shell
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_api.py tests/test_abi.py -m gpu -q 2>&1 | grep -v "Warning: Particle" > gpurun_out/r3a_tests.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r3a_tests.log | head -60
