# round 2 (session 3), job 4k: multi-device engine with one worker thread per device — tests (slabs on one GPU), host shell, sanitizer
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_api.py -m gpu -q 2>&1 | grep -v "Warning: Particle" | grep "^E  \|^FAILED\|passed\|failed" | head
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "dam_break or adaptive or rebalances" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -3
timeout 600 python tools/bench_multi.py --devices 0,0,0,0 --scene dam_break_1M --scaling strong --steps 50 --warmup 20 --check 2>&1 | tail -1 | cut -c1-900
