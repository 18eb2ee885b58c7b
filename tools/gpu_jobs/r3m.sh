# round 2 (session 3), job m: the halo-sliver regression test passes with the fix and FAILS on a build without it; rebalancing test
set -x
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -8
SPHB_LIB=$PWD/tune/libsphb_nosliver.so timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k sliver 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -5
