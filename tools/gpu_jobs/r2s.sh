timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_staged_equals_global and micro_pair-4" 2>&1 | grep -E "Error|error|assert" | head -10
SPHB_PAIR_MODE=1 timeout 120 python bench.py --no-cpu --warmup 2 --steps 2 2>&1 | tail -5 | cut -c1-400
SPHB_PAIR_MODE=1 timeout 300 compute-sanitizer --tool memcheck python bench.py --no-cpu --warmup 1 --steps 1 --scene dam_break_13k 2>&1 | grep -E "Invalid|ERROR|at 0x|by thread|Address" | head -20
