# round 2 (session 3), job k: halo sliver — regression test, 2-slab vs single at 2 x 10.6 M over 90 steps, slab + multi suites
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_slab_gpu.py -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head
timeout 1200 python tools/debug/multi_bisect.py dam_break_10M 0,0 90 2>&1 | tail -10
# the regression test must FAIL on a build without the sliver
SPHB_LIB=$PWD/tune/libsphb_nosliver.so timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k sliver 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -5
