# round 2 (session 3), job o: does the slab-vs-single difference (build without the halo sliver) depend on the staged density pass?
set -x
SPHB_PAIR_MODE=0 SPHB_LIB=$PWD/tune/libsphb_nosliver.so timeout 1500 python tools/debug/multi_bisect2.py 46 50 2>&1 | tail -12
SPHB_PAIR_MODE=1 SPHB_LIB=$PWD/tune/libsphb_nosliver.so timeout 1500 python tools/debug/multi_bisect2.py 46 50 2>&1 | tail -12
