set -x
timeout 600 python tools/debug/e2e_overlap.py 2>&1 | tail -9
