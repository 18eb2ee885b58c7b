# round 2 (session 3), job j: does a third halo layer remove the slab-vs-single difference that starts with the first migration?
set -x
LAYERS=3 timeout 1200 python tools/debug/multi_bisect.py dam_break_10M 0,0 90 2>&1 | tail -12
