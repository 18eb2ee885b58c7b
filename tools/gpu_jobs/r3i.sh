# round 2 (session 3), job i: are force-pass records of halo particles without a local density ever read? (NaN-poisoned build)
set -x
SPHB_LIB=$PWD/tune/libsphb_poison.so timeout 1200 python tools/debug/multi_bisect.py dam_break_10M 0,0 60 2>&1 | tail -12
