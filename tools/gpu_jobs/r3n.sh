# round 2 (session 3), job n: first step and first particles at which 2 slabs WITHOUT the halo sliver leave the single-context run
set -x
SPHB_LIB=$PWD/tune/libsphb_nosliver.so timeout 1500 python tools/debug/multi_bisect2.py 46 60 2>&1 | tail -30
